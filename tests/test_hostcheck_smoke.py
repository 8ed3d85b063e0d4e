"""The smoke kernels' bodies (csrc/smoke_ops.cuh: monotonic Catmull-Rom, the cubic sampler, the adaptive mid-point
back-trace, scalar / face advection, buoyancy) executed on the CPU through tests/hostcheck, against the C oracle
(bit-identical to the unmodified reference: tests/test_oracle_smoke.py) — bit for bit."""
import ctypes as C

import numpy as np
import pytest

from oracle_api import Grid, MgpcgParams, SmokeParams, have_port, load_port
from test_hostcheck_collider import _load_hostcheck, dp, exact
from test_oracle_smoke import EMIT_GRIDS, SMOKE_COLLIDERS, plume, source_shapes

pytestmark = pytest.mark.skipif(not have_port(), reason="oracle/libjetport.so not built")


@pytest.fixture(scope="module")
def hc():
    lib = _load_hostcheck()
    if lib is None:
        pytest.skip("nvcc not available to build tests/hostcheck")
    return lib


@pytest.fixture(scope="module")
def port():
    o = load_port()
    o.set_num_threads(0)
    return o


def developed(port, n, domain, levels, gravity, steps=3, advection="cubic"):
    prm = SmokeParams.make(Grid.cube(n, domain), mgpcg=MgpcgParams.make(max_iterations=60, max_levels=levels), gravity=gravity,
                           advection=advection)
    s = port.smoke(prm)
    den, temp, u, v, w = plume(n, domain, seed=3)
    s.set_fields(den, temp, 0 * u, 0 * v, 0 * w)
    for _ in range(steps):
        s.step(1.0 / 30.0)
    return prm, s


@pytest.mark.parametrize("n,domain,levels,gravity", [(16, 1.0, 3, (0.0, -9.8, 0.0)), (12, 0.9, 2, (3.0, -4.0, 1.0))])
def test_buoyancy(hc, port, n, domain, levels, gravity):
    prm, s = developed(port, n, domain, levels, gravity)
    f = s.get_fields()
    dt = 1.0 / 60.0
    s.compute_external_forces(dt)
    want = s.get_fields()
    t_sum = 0.0
    for x in f["temperature"].ravel():      # the reference's serial k, j, i sum
        t_sum += float(x)
    u, v, w = f["u"].copy(), f["v"].copy(), f["w"].copy()
    assert hc.hc_smoke_buoyancy(C.byref(prm.grid), dp(f["density"]), dp(f["temperature"]), C.c_double(prm.buoyancy_density_factor),
                                C.c_double(prm.buoyancy_temperature_factor), C.c_double(t_sum), (C.c_double * 3)(*gravity),
                                C.c_double(dt), C.c_int(prm.closed_domain_boundary_flag), dp(u), dp(v), dp(w)) == 0
    for a, k in zip((u, v, w), "uvw"):
        exact(a, want[k], "buoyancy " + k)
    assert not np.array_equal(want["v"], f["v"])
    s.close()


@pytest.mark.parametrize("advection", ["cubic", "linear"])
@pytest.mark.parametrize("n,domain,levels,gravity,dt", [(16, 1.0, 3, (0.0, -9.8, 0.0), 1.0 / 30.0), (12, 0.9, 2, (3.0, -4.0, 1.0), 1.0 / 15.0)])
def test_advection(hc, port, n, domain, levels, gravity, dt, advection):
    prm, s = developed(port, n, domain, levels, gravity, steps=5, advection=advection)
    f = s.get_fields()
    s.compute_advection(dt)
    want = s.get_fields()
    g = {k: f[k].copy() for k in ("density", "temperature", "u", "v", "w")}
    assert hc.hc_smoke_advect(C.byref(prm.grid), C.c_double(dt), C.c_int(prm.closed_domain_boundary_flag), C.c_int(prm.advection), dp(g["density"]),
                              dp(g["temperature"]), dp(g["u"]), dp(g["v"]), dp(g["w"])) == 0
    for k in g:
        exact(g[k], want[k], "advected " + k)
    assert np.abs(want["density"] - f["density"]).max() > 1e-3   # the plume really moved
    s.close()


@pytest.mark.parametrize("which", range(4))
@pytest.mark.parametrize("gi", range(len(EMIT_GRIDS)))
def test_emit_step(hc, port, gi, which):
    from oracle_api import ColliderShape
    grid = EMIT_GRIDS[gi]
    prm = SmokeParams.make(grid, mgpcg=MgpcgParams.make(max_iterations=40, max_levels=1))
    s = port.smoke(prm)
    rng = np.random.default_rng(10 * gi + which)
    nx, ny, nz = grid.res
    den0, t0 = rng.uniform(0, 0.5, (nz, ny, nx)), rng.uniform(0, 0.5, (nz, ny, nx))
    s.set_fields(den0, t0, np.zeros((nz, ny, nx + 1)), np.zeros((nz, ny + 1, nx)), np.zeros((nz + 1, ny, nx)))
    shapes = source_shapes(grid, which)
    s.emit_step(shapes, (0.0, 1.0), (0.25, 2.0))
    want = s.get_fields()
    arr = (ColliderShape * len(shapes))()
    for i, sh in enumerate(shapes):
        arr[i].kind = {"box": 1, "sphere": 2, "cylinder": 3, "plane": 4}[sh["kind"]]
        arr[i].a = (C.c_double * 3)(*sh["a"])
        b = tuple(sh.get("b", (0.0, 0.0, 0.0))) + (0.0, 0.0, 0.0)
        arr[i].b = (C.c_double * 3)(*b[:3])
    den, temp = den0.copy(), t0.copy()
    args = [C.byref(grid), arr, C.c_uint32(len(shapes)), C.c_int(int(len(shapes) != 1))]
    assert hc.hc_smoke_emit_step(*args, C.c_double(0.0), C.c_double(1.0), dp(den)) == 0
    assert hc.hc_smoke_emit_step(*args, C.c_double(0.25), C.c_double(2.0), dp(temp)) == 0
    exact(den, want["density"], "emitted density"); exact(temp, want["temperature"], "emitted temperature")
    s.close()


@pytest.mark.parametrize("ci", range(len(SMOKE_COLLIDERS)))
def test_advection_with_collider(hc, port, ci):
    """computeAdvection with a rigid collider on a developed plume: boundary handling in the back-trace, untouched cells inside the
    collider, extrapolateIntoCollider of the scalars, the collider boundary condition on the advected velocity."""
    from oracle_api import ColliderShape
    grid, levels, shapes, kw, advection = SMOKE_COLLIDERS[ci]
    prm = SmokeParams.make(grid, mgpcg=MgpcgParams.make(max_iterations=60, max_levels=levels), advection=advection)
    s = port.smoke(prm)
    nx, ny, nz = grid.res
    lo = np.array([grid.ox, grid.oy, grid.oz]); size = np.array([nx * grid.hx, ny * grid.hy, nz * grid.hz])
    source = [dict(kind="box", a=tuple(lo + np.array([0.35, -1.0, 0.35]) * size), b=tuple(lo + np.array([0.65, 0.15, 0.65]) * size))]
    s.set_collider_shapes(shapes, **kw)
    for _ in range(8):
        s.emit_step(source, (0.0, 1.0), (0.0, 1.0))
        s.step(1.0 / 30.0)
    f = s.get_fields()
    dt = 1.0 / 20.0
    s.compute_advection(dt)
    want = s.get_fields()
    arr = (ColliderShape * len(shapes))()
    for i, sh in enumerate(shapes):
        arr[i].kind = {"box": 1, "sphere": 2, "cylinder": 3, "plane": 4}[sh["kind"]]
        arr[i].a = (C.c_double * 3)(*sh["a"])
        b = tuple(sh.get("b", (0.0, 0.0, 0.0))) + (0.0, 0.0, 0.0)
        arr[i].b = (C.c_double * 3)(*b[:3])
    v3 = lambda k: (C.c_double * 3)(*kw.get(k, (0.0, 0.0, 0.0)))
    g = {k: f[k].copy() for k in ("density", "temperature", "u", "v", "w")}
    assert hc.hc_smoke_advect_collider(C.byref(prm.grid), C.c_double(dt), C.c_int(prm.closed_domain_boundary_flag), C.c_int(prm.advection),
                                       arr, C.c_uint32(len(shapes)), C.c_int(int(len(shapes) > 1)), v3("linear_velocity"),
                                       v3("angular_velocity"), C.c_double(kw.get("friction", 0.0)), C.c_double(prm.max_cfl),
                                       dp(g["density"]), dp(g["temperature"]), dp(g["u"]), dp(g["v"]), dp(g["w"])) == 0
    for k in g:
        exact(g[k], want[k], "advected " + k)
    s.close()


def test_cfl(hc, port):
    prm, s = developed(port, 16, 1.0, 3, (3.0, -4.0, 1.0), steps=6)
    f = s.get_fields()
    for dt in (1.0 / 60.0, 0.25, 2.0):
        got = C.c_double(0.0)
        assert hc.hc_smoke_cfl(C.byref(prm.grid), dp(f["u"]), dp(f["v"]), dp(f["w"]), (C.c_double * 3)(3.0, -4.0, 1.0), C.c_double(dt),
                               C.byref(got)) == 0
        assert got.value == s.cfl(dt)
    s.close()

