"""GPU parity for the device-resident GridSmokeSolver3 (BASELINE config 5).

csrc/smoke.cu is checked against the pinned oracle (tests/test_oracle_smoke.py: port == unmodified reference); its kernel
bodies are also verified on the CPU (tests/test_hostcheck_smoke.py).  Green on B200 since round 2's first GPU call.

Expected: advection bit-exact on identical inputs (per-cell arithmetic); buoyancy within 1e-12 (the ambient temperature
is a tree sum on the GPU, a serial sum in the reference); pressure <= 1e-6 relative where the oracle converged; fields
after 5 steps within 1e-9 of the field scale."""
import numpy as np
import pytest

import golden_cases as gc
from gpu_common import checker, jetcu_or_fail
from oracle_api import Grid, MgpcgParams, SmokeParams
from test_oracle_smoke import EMIT_GRIDS, SMOKE_COLLIDERS, plume, smoke_collider_run, source_shapes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jc():
    return jetcu_or_fail()


@pytest.fixture(scope="module")
def oracle():
    return checker()


def close(a, b, tol, what):
    scale = max(float(np.abs(b).max()), 1e-300)
    err = float(np.abs(a - b).max()) / scale
    assert err <= tol, "%s: relative L-inf error %.3e > %.1e" % (what, err, tol)


@pytest.mark.parametrize("advection", ["cubic", "linear"])     # CubicSemiLagrangian3 (Jet's default) / SemiLagrangian3
@pytest.mark.parametrize("n,domain,levels,gravity", [(16, 1.0, 3, (0.0, -9.8, 0.0)), (12, 0.9, 2, (3.0, -4.0, 1.0)), (32, 1.0, 4, (0.0, -9.8, 0.0))])
def test_smoke_stages(jc, oracle, n, domain, levels, gravity, advection):
    prm = SmokeParams.make(Grid.cube(n, domain), mgpcg=MgpcgParams.make(max_iterations=60, max_levels=levels), gravity=gravity,
                           advection=advection)
    a, b = jc.smoke(prm), oracle.smoke(prm)
    den, temp, u, v, w = plume(n, domain, seed=3)
    b.set_fields(den, temp, 0 * u, 0 * v, 0 * w)
    for _ in range(3):
        b.step(1.0 / 30.0)            # a developed plume from the oracle: identical inputs for every stage below
    dt = 1.0 / 30.0
    f = b.get_fields()
    sync = lambda: a.set_fields(*(b.get_fields()[k] for k in ("density", "temperature", "u", "v", "w")))
    sync()
    a.compute_external_forces(dt); b.compute_external_forces(dt)
    fa, fb = a.get_fields(), b.get_fields()
    vs = max(float(np.abs(fb[k]).max()) for k in "uvw")
    for k in "uvw":
        assert np.abs(fa[k] - fb[k]).max() <= 1e-12 * vs, "buoyancy " + k
    sync()
    ra, rb = a.compute_pressure(dt), b.compute_pressure(dt)
    fa, fb = a.get_fields(), b.get_fields()
    if rb.true_residual_l2 <= 1e-8 * max(rb.b_l2, 1e-300):
        close(fa["pressure"], fb["pressure"], 1e-6, "pressure")
        for k in "uvw":
            assert np.abs(fa[k] - fb[k]).max() <= 1e-6 * vs, "projected " + k
    sync()
    a.compute_advection(dt); b.compute_advection(dt)
    fa, fb = a.get_fields(), b.get_fields()
    for k in ("density", "temperature", "u", "v", "w"):
        assert np.array_equal(fa[k], fb[k]), "advected %s: max abs diff %g" % (k, np.abs(fa[k] - fb[k]).max())
    sync()
    a.end_advance_time_step(dt); b.end_advance_time_step(dt)
    fa, fb = a.get_fields(), b.get_fields()
    assert np.array_equal(fa["density"], fb["density"]) and np.array_equal(fa["temperature"], fb["temperature"])
    a.close(); b.close()


@pytest.mark.parametrize("n,levels", [(16, 3), (32, 4)])
def test_smoke_steps(jc, oracle, n, levels):
    prm = SmokeParams.make(Grid.cube(n), mgpcg=MgpcgParams.make(max_iterations=60, max_levels=levels))
    a, b = jc.smoke(prm), oracle.smoke(prm)
    den, temp, u, v, w = plume(n, seed=3)
    a.set_fields(den, temp, 0 * u, 0 * v, 0 * w); b.set_fields(den, temp, 0 * u, 0 * v, 0 * w)
    for s in range(5):
        ra, rb = a.step(1.0 / 30.0), b.step(1.0 / 30.0)
    fa, fb = a.get_fields(), b.get_fields()
    vs = max(float(np.abs(fb[k]).max()) for k in "uvw")
    for k in "uvw":
        assert np.abs(fa[k] - fb[k]).max() <= 1e-9 * vs, "%s after 5 steps: %g" % (k, np.abs(fa[k] - fb[k]).max() / vs)
    close(fa["density"], fb["density"], 1e-9, "density after 5 steps")
    close(fa["temperature"], fb["temperature"], 1e-9, "temperature after 5 steps")
    a.close(); b.close()


def test_golden_smoke(jc):
    gc.smoke(jc.smoke, exact_bits=False)


@pytest.mark.parametrize("which", range(4))
@pytest.mark.parametrize("gi", range(len(EMIT_GRIDS)))
def test_emit_step(jc, oracle, gi, which):
    """Device-side VolumeGridEmitter3 step-function source; sin() is device libm: 1e-15 instead of bit-exact."""
    grid = EMIT_GRIDS[gi]
    prm = SmokeParams.make(grid, mgpcg=MgpcgParams.make(max_iterations=40, max_levels=1))
    a, b = jc.smoke(prm), oracle.smoke(prm)
    rng = np.random.default_rng(10 * gi + which)
    nx, ny, nz = grid.res
    den0, t0 = rng.uniform(0, 0.5, (nz, ny, nx)), rng.uniform(0, 0.5, (nz, ny, nx))
    for s in (a, b):
        s.set_fields(den0, t0, np.zeros((nz, ny, nx + 1)), np.zeros((nz, ny + 1, nx)), np.zeros((nz + 1, ny, nx)))
        s.emit_step(source_shapes(grid, which), (0.0, 1.0), (0.25, 2.0))
    fa, fb = a.get_fields(), b.get_fields()
    for k in ("density", "temperature"):
        assert np.abs(fa[k] - fb[k]).max() <= 1e-14, k
    a.close(); b.close()


@pytest.mark.parametrize("ci", range(len(SMOKE_COLLIDERS)))
def test_smoke_with_collider(jc, oracle, ci):
    """Rising smoke hitting a rigid collider (src/examples/smoke_sim example 1), 10 steps with a continuous emitter: within 1e-9
    of the field scale (the ambient temperature is a tree sum on the GPU; sin / atan2 are device libm)."""
    fa, fb = smoke_collider_run(jc, SMOKE_COLLIDERS[ci], 10), smoke_collider_run(oracle, SMOKE_COLLIDERS[ci], 10)
    for k in ("density", "temperature", "u", "v", "w"):
        close(fa[-1][k], fb[-1][k], 1e-9 if k in ("density", "temperature") else 1e-7, "after 10 steps: " + k)


def test_cfl_and_sub_steps(jc, oracle):
    """GridFluidSolver3::cfl / numberOfSubTimeSteps from the device-resident velocity (max is exact: bit-identical)."""
    prm = SmokeParams.make(Grid.cube(16), mgpcg=MgpcgParams.make(max_iterations=60, max_levels=3), max_cfl=0.5)
    a, b = jc.smoke(prm), oracle.smoke(prm)
    den, temp, u, v, w = plume(16, 1.0, seed=3)
    for s in (a, b):
        s.set_fields(den, temp, u, v, w)
    for dt in (1.0 / 60.0, 0.25, 2.0):
        assert a.cfl(dt) == b.cfl(dt) and a.num_sub_steps(dt) == b.num_sub_steps(dt)
    a.close(); b.close()


@pytest.mark.parametrize("advection", ["cubic", "linear"])
def test_golden_smoke_example1(jc, advection):
    gc.smoke_example1(jc.smoke, exact_bits=False, advection=advection)

