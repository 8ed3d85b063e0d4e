"""GPU vs oracle on RANDOM configurations (the cases of tests/test_oracle_fuzz.py, where the oracle is pinned to the
unmodified reference bit for bit): non-zero origins, non-cubic and anisotropic grids, odd wall flags, tilted gravity, every
transfer mode, one or two MG levels, random analytic colliders.  Two sub-steps; BASELINE tolerances: positions <= 1e-5 of the
domain size, pressure <= 1e-6 relative where the oracle's own solve converged; particle SDF and markers bit-exact on step 0
when no cylinder (device libm) is involved.  Every fourth configuration smooths with the natural Gauss-Seidel ordering (the
reference constructors' default)."""
import numpy as np
import pytest

from gpu_common import checker, jetcu_or_fail
from oracle_api import FLIP, Grid, HybridParams, MgpcgParams
from test_oracle_fuzz import NONCUBIC, random_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jc():
    return jetcu_or_fail()


@pytest.fixture(scope="module")
def oracle():
    return checker()


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(jc, oracle, seed):
    prm, pos, vel, collider = random_case(np.random.default_rng(1000 + seed), natural=seed % 4 == 3)   # every 4th: natural ordering
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    for h in (a, b):
        if collider:
            h.set_collider_shape(**collider)
        h.set_particles(pos, vel)
    g = prm.grid
    L = max(g.nx * g.hx, g.ny * g.hy, g.nz * g.hz)
    libm = bool(collider) and collider["kind"] == "cylinder"
    for step in range(2):
        sa, sb = a.step(1.0 / 60.0), b.step(1.0 / 60.0)
        pa, va = a.get_particles(); pb, vb = b.get_particles()
        assert np.abs(pa - pb).max() <= 1e-5 * L, "positions step %d: %g" % (step, np.abs(pa - pb).max())
        assert np.abs(va - vb).max() <= 1e-6 * max(np.abs(vb).max(), 1e-300) + 1e-9, "velocities step %d" % step
        if sb.pcg_true_residual_l2 <= 1e-8 * max(sb.pcg_b_l2, 1e-300):
            pr_a, pr_b = a.get_pressure(), b.get_pressure()
            assert np.abs(pr_a - pr_b).max() <= 1e-6 * max(np.abs(pr_b).max(), 1e-300), "pressure step %d" % step
        if step == 0 and not libm:
            assert np.array_equal(a.get_sdf(), b.get_sdf()), "particle sdf"
            for x, y, nm in zip(a.get_markers(), b.get_markers(), "uvw"):
                assert np.array_equal(x, y), "marker " + nm
    a.close(); b.close()


@pytest.mark.parametrize("shape", NONCUBIC)
def test_noncubic_multilevel_extension(jc, oracle, shape):
    """ny != nz with three MG levels (no reference answer: see tests/test_oracle_fuzz.py): CUDA vs the C oracle, which clamp
    the prolongation against n.z alike — the usual tolerances."""
    nx, ny, nz = shape
    h = 1.0 / nx
    prm = HybridParams.make(Grid(nx, ny, nz, h, h, h, 0.0, 0.0, 0.0), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=3))
    from oracle_api import load_port
    port = load_port(); port.set_num_threads(0)          # the reference library itself is undefined on these shapes
    a, b = jc.hybrid(prm), port.hybrid(prm)
    for s in (a, b):
        s.emit_box((0.0, 0.0, 0.0), (0.4 * nx * h, 0.6 * ny * h, nz * h), 0.5 * h, (0.5, 0.0, 0.0))
    for step in range(2):
        sa, sb = a.step(1.0 / 60.0), b.step(1.0 / 60.0)
        assert sa.pcg_iterations == sb.pcg_iterations
        pa, pb = a.get_particles()[0], b.get_particles()[0]
        assert np.abs(pa - pb).max() <= 1e-5 * max(nx, ny, nz) * h
    a.close(); b.close()

