"""GPU parity, SURVEY §8f-4: GridBackwardEulerDiffusionSolver3 (stand-alone, through the C ABI) and GridFluidSolver3::
computeViscosity inside the hybrid step, against the oracle (the unmodified reference where its shim is built).  Markers are
exact; the solves go through ICCG, whose dot products the device reduces with a fixed tree (the reference serially), so the
fields are compared within 1e-9 relative — far inside north_star's 1e-6 for solver outputs."""
import numpy as np
import pytest

from gpu_common import checker, jetcu_or_fail
from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams
from test_oracle_diffusion import fields

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jc():
    return jetcu_or_fail()


@pytest.fixture(scope="module")
def oracle():
    return checker()


def close(a, b, rtol, what):
    scale = max(float(np.abs(b).max()), 1e-300)
    err = float(np.abs(a - b).max()) / scale
    assert err <= rtol, "%s: relative L-inf %g > %g" % (what, err, rtol)


@pytest.mark.parametrize("boundary_type", [0, 1])
@pytest.mark.parametrize("n,shape3,with_collider,with_fluid",
                         [(12, (12, 14, 10), True, True), (10, (10, 10, 10), False, True), (9, (9, 8, 11), True, False),
                          (32, (32, 32, 32), True, True)])
def test_face_solve(jc, oracle, n, shape3, with_collider, with_fluid, boundary_type):
    grid, u, v, w, col, flu = fields(n, 7 + n, shape3, with_collider, with_fluid)
    coef, dt = 0.05, 1.0 / 60.0
    a = jc.diffusion_solve_faces(grid, u, v, w, coef, dt, boundary_type, col, flu)
    b = oracle.diffusion_solve_faces(grid, u, v, w, coef, dt, boundary_type, col, flu)
    for q, name in enumerate("uvw"):
        close(a[q], b[q], 1e-9, name)
    # the w solve is the one both sides report: same iteration count up to the rounding of the dots
    assert abs(int(a[3][2].iterations) - int(b[3][2].iterations)) <= max(1, int(b[3][2].iterations) // 20)


@pytest.mark.parametrize("boundary_type", [0, 1])
def test_cell_solve_and_markers(jc, oracle, boundary_type):
    grid, u, v, w, col, flu = fields(12, 3, (12, 14, 10), True, True)
    src = np.random.default_rng(5).standard_normal((10, 14, 12))
    a = jc.diffusion_solve_cells(grid, src, 0.2, 0.01, boundary_type, col, flu, want_markers=True)
    b = oracle.diffusion_solve_cells(grid, src, 0.2, 0.01, boundary_type, col, flu, want_markers=True)
    assert np.array_equal(a[2], b[2]), "markers differ in %d samples" % int((a[2] != b[2]).sum())
    close(a[0], b[0], 1e-9, "diffused scalar")


def test_diffusion_conserves_constant_and_smooths(jc):
    """Size-independent properties at 64^3, all fluid, Neumann: a constant stays constant; a random field keeps its mean (the
    system is I - c L with zero-flux boundaries) and loses variance."""
    n = 64
    grid = Grid.cube(n)
    const = np.full((n, n, n), 3.25)
    out, res = jc.diffusion_solve_cells(grid, const, 0.01, 0.01, 1)
    assert np.abs(out - 3.25).max() <= 1e-12
    src = np.random.default_rng(1).standard_normal((n, n, n))
    out, res = jc.diffusion_solve_cells(grid, src, 1e-4, 0.01, 1)
    assert abs(out.mean() - src.mean()) <= 1e-10
    assert out.var() < 0.97 * src.var()      # c = dt * coefficient / h^2 = 4.1e-3: first B200 run 0.9537


@pytest.mark.parametrize("mode", [PIC, FLIP, APIC])
def test_hybrid_steps_with_viscosity(jc, oracle, mode):
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=60, max_levels=3))
    hs = [lib.hybrid(prm) for lib in (jc, oracle)]
    for h in hs:
        h.emit_box((0.0, 0.0, 0.0), (0.5, 0.7, 1.0), 0.5 / n, (0.4, -0.1, 0.2))
        h.set_collider_shape("box", (0.6, 0.0, 0.3), (0.8, 0.3, 0.7))
        h.set_viscosity(0.08)
    for step in range(4):
        for h in hs:
            h.step(1.0 / 120.0)
        for a, b, name in zip(hs[0].get_velocity(), hs[1].get_velocity(), "uvw"):
            assert np.abs(a - b).max() <= 1e-7, "step %d %s: %g" % (step, name, np.abs(a - b).max())
    (pa, va), (pb, vb) = hs[0].get_particles(), hs[1].get_particles()
    assert np.abs(pa - pb).max() <= 1e-5 and np.abs(va - vb).max() <= 1e-5
    # the viscous solve is really part of the step
    r = hs[0].compute_viscosity(1.0 / 120.0)
    assert all(int(x.iterations) > 0 for x in r)
    for h in hs:
        h.close()


@pytest.mark.parametrize("ci,diffusion", [(0, (0.02, 0.001, 0.003)), (2, (0.0, 0.002, 0.0)), (4, (0.05, 0.0, 0.004))])
def test_smoke_with_viscosity_and_diffusion(jc, oracle, ci, diffusion):
    """GridFluidSolver3::computeViscosity and GridSmokeSolver3::computeDiffusion's scalar solves in the smoke step, next to a
    rigid collider (the scenes of tests/test_oracle_smoke.py, where port == unmodified reference bit for bit)."""
    from test_oracle_smoke import SMOKE_COLLIDERS, smoke_collider_run
    fa = smoke_collider_run(jc, SMOKE_COLLIDERS[ci], 4, diffusion)
    fb = smoke_collider_run(oracle, SMOKE_COLLIDERS[ci], 4, diffusion)
    for step, (a, b) in enumerate(zip(fa, fb)):
        for k in ("density", "temperature", "u", "v", "w"):
            close(a[k], b[k], 1e-6, "step %d %s" % (step, k))
