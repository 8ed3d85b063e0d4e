"""GridSmokeSolver3 (BASELINE config 5): the C restatement (oracle/jet_port_smoke.inc) against the
UNMODIFIED reference — buoyancy, all-liquid pressure projection with MGPCG, cubic semi-Lagrangian advection of density /
temperature / velocity, decay.  Every stage is per-cell / per-face arithmetic (the ambient-temperature sum is serial in
both), so the port must be BIT-IDENTICAL.  CPU only."""
import numpy as np
import pytest

from oracle_api import Grid, MgpcgParams, SmokeParams, have_port, have_ref, load_port, load_ref

pytestmark = pytest.mark.skipif(not (have_port() and have_ref()), reason="needs both oracle libraries")


@pytest.fixture(scope="module")
def libs():
    ref, port = load_ref(), load_port()
    ref.set_num_threads(0)
    port.set_num_threads(0)
    return ref, port


def plume(n, domain=1.0, seed=0):
    """Warm dense blob near the floor plus a rough divergent velocity field."""
    h = domain / n
    c = (np.arange(n) + 0.5) * h
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
    r = np.sqrt((X - 0.5 * domain) ** 2 + (Y - 0.25 * domain) ** 2 + (Z - 0.45 * domain) ** 2)
    den = np.clip(1.0 - r / (0.2 * domain), 0.0, 1.0)
    temp = 2.0 * den + 0.1
    rng = np.random.default_rng(seed)
    u = 0.3 * rng.uniform(-1, 1, (n, n, n + 1)); v = 0.3 * rng.uniform(-1, 1, (n, n + 1, n)); w = 0.3 * rng.uniform(-1, 1, (n + 1, n, n))
    return np.ascontiguousarray(den), np.ascontiguousarray(temp), u, v, w


def same(a, b, what):
    for k in a:
        assert np.array_equal(a[k], b[k]), "%s: %s max abs diff %g" % (what, k, float(np.abs(a[k] - b[k]).max()))


@pytest.mark.parametrize("n,domain,levels", [(16, 1.0, 3), (12, 0.9, 2)])
def test_smoke_stages_bit_exact(libs, n, domain, levels):
    ref, port = libs
    prm = SmokeParams.make(Grid.cube(n, domain), mgpcg=MgpcgParams.make(max_iterations=60, max_levels=levels))
    a, b = ref.smoke(prm), port.smoke(prm)
    fields = plume(n, domain)
    a.set_fields(*fields); b.set_fields(*fields)
    dt = 1.0 / 60.0
    a.step(dt); b.step(dt)                      # also brings the reference's boundary-condition solver up
    same(a.get_fields(), b.get_fields(), "after one step")
    a.compute_external_forces(dt); b.compute_external_forces(dt)
    same(a.get_fields(), b.get_fields(), "buoyancy")
    ra, rb = a.compute_pressure(dt), b.compute_pressure(dt)
    assert ra.iterations == rb.iterations and ra.last_residual == rb.last_residual
    same(a.get_fields(), b.get_fields(), "pressure")
    a.compute_advection(dt); b.compute_advection(dt)
    same(a.get_fields(), b.get_fields(), "advection")
    a.end_advance_time_step(dt); b.end_advance_time_step(dt)
    same(a.get_fields(), b.get_fields(), "decay")
    a.close(); b.close()


@pytest.mark.parametrize("n,levels,gravity", [(16, 3, (0.0, -9.8, 0.0)), (16, 3, (3.0, -4.0, 1.0)), (12, 2, (0.0, -9.8, 0.0)),
                                              (12, 2, (0.0, 0.0, 0.0))])
def test_smoke_steps_bit_exact(libs, n, levels, gravity):
    """A plume rising from rest in a closed box (the reference's MGPCG converges here: 9-26 iterations per step; with an
    open wall or a rough initial velocity it diverges and the back-trace then takes thousands of sub-steps)."""
    ref, port = libs
    prm = SmokeParams.make(Grid.cube(n), mgpcg=MgpcgParams.make(max_iterations=60, max_levels=levels), gravity=gravity,
                           density_decay=0.01, temperature_decay=0.02)
    a, b = ref.smoke(prm), port.smoke(prm)
    den, temp, u, v, w = plume(n, seed=3)
    a.set_fields(den, temp, 0 * u, 0 * v, 0 * w); b.set_fields(den, temp, 0 * u, 0 * v, 0 * w)
    for s in range(5):
        ra, rb = a.step(1.0 / 30.0), b.step(1.0 / 30.0)
        assert ra.iterations == rb.iterations, (s, ra.iterations, rb.iterations)
        assert rb.true_residual_l2 <= 1e-6 * max(rb.b_l2, 1e-300) or rb.iterations < 60
        same(a.get_fields(), b.get_fields(), "step %d" % s)
    fb = b.get_fields()
    vmax = max(np.abs(fb[k]).max() for k in "uvw")
    assert 0.05 < vmax < 10.0 and fb["density"].max() > 0.1
    a.close(); b.close()
