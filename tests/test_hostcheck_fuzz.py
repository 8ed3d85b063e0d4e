"""The CUDA kernel bodies (tests/hostcheck: the product's __host__ __device__ source behind plain loops) against the oracle on
RANDOM grids — non-zero origins, non-cubic shapes, anisotropic and non-power-of-two spacings, odd wall flags: P2G, G2P (PIC /
APIC), midpoint advection with wall clamping, the collider boundary condition.  The fixed hostcheck cases all sit on
origin-0 cubes; index arithmetic that silently assumes that would pass them and fail here."""
import ctypes as C

import numpy as np
import pytest

from oracle_api import APIC, PIC, ColliderShape, Grid, HybridParams, MgpcgParams, have_port, load_port
from test_hostcheck_collider import _load_hostcheck, dp, exact
from test_hostcheck_p2g import run_hc

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


def random_grid(rng):
    nyz = int(rng.choice([6, 8, 10, 12]))
    nx = int(rng.choice([nyz, nyz + 2, nyz + 3, 2 * nyz]))
    h = float(rng.choice([1.0 / nyz, 0.07, 0.125, 0.1]))
    aniso = rng.random() < 0.4
    hy, hz = h * (1.25 if aniso else 1.0), h * (0.8 if aniso else 1.0)
    origin = tuple(float(x) for x in rng.choice([0.0, 0.3, -0.45, 1.5], 3))
    return Grid(nx, nyz, nyz, h, hy, hz, *origin)


def pow2(x):
    m, _ = np.frexp(x)
    return m == 0.5


def variants(grid):
    return [0, 2] + ([1, 3] if pow2(grid.hx) and pow2(grid.hy) and pow2(grid.hz) else [])


def populated(port, grid, mode, rng, flag=63):
    prm = HybridParams.make(grid, mode=mode, mgpcg=MgpcgParams.make(max_iterations=30, max_levels=1),
                            closed_domain_boundary_flag=flag, max_cfl=float(rng.choice([5.0, 2.5])))
    h = port.hybrid(prm)
    lo = np.array([grid.ox, grid.oy, grid.oz]); size = np.array([grid.nx * grid.hx, grid.ny * grid.hy, grid.nz * grid.hz])
    n = int(rng.integers(500, 4000))
    pos = lo + rng.uniform(0.0, 0.75, (n, 3)) * size
    pos[:20] = lo + size * rng.choice([0.0, 1.0], (20, 3))          # corners and faces of the domain
    vel = rng.uniform(-1.0, 1.0, (n, 3)) * 2.0
    h.set_particles(np.ascontiguousarray(pos), vel)
    return prm, h


@pytest.mark.parametrize("seed", range(10))
def test_p2g_random_grid(hc, port, seed):
    rng = np.random.default_rng(2000 + seed)
    grid = random_grid(rng)
    mode = APIC if seed % 2 else PIC
    prm, h = populated(port, grid, mode, rng)
    aff = None
    if mode == APIC:
        h.step(1.0 / 120.0)
        aff = h.get_affine()
    pos, vel = h.get_particles()
    h.transfer_particles_to_grids()
    want, want_m = h.get_velocity(), h.get_markers()
    scale = max(float(np.abs(a).max()) for a in want)
    for tz in (0, 4 if mode == APIC else 8):     # 0: the staggered-lattice kernel (default), 4 / 8: the window kernel
        got, got_m = run_hc(hc, grid, mode == APIC, tz, 1, pos, vel, aff)
        for a, b, am, bm, nm in zip(got, want, got_m, want_m, "uvw"):
            exact(am, bm, "marker " + nm)
            assert np.abs(a - b).max() <= 1e-12 * scale, "tz %d %s: %g" % (tz, nm, np.abs(a - b).max() / scale)
    h.close()


@pytest.mark.parametrize("seed", range(10))
def test_g2p_and_move_random_grid(hc, port, seed):
    rng = np.random.default_rng(3000 + seed)
    grid = random_grid(rng)
    mode = APIC if seed % 2 else PIC
    flag = int(rng.choice([63, 0b101011, 0b111100, 0]))
    prm, h = populated(port, grid, mode, rng, flag)
    dt = 1.0 / 60.0
    h.step(dt)
    u, v, w = h.get_velocity()
    pos, vel = h.get_particles()
    h.transfer_grids_to_particles()
    _, want_v = h.get_particles()
    if mode == PIC:
        for var in variants(grid):
            got = np.zeros_like(want_v)
            assert hc.hc_g2p_pic(C.byref(grid), dp(u), dp(v), dp(w), C.c_int(var), C.c_uint64(len(pos)), dp(pos), dp(got)) == 0
            exact(got, want_v, "PIC G2P (variant %d)" % var)
    else:
        want_c = h.get_affine()
        gv = np.zeros_like(want_v)
        gc = [np.zeros_like(want_v) for _ in range(3)]
        assert hc.hc_g2p_apic(C.byref(grid), dp(u), dp(v), dp(w), C.c_uint64(len(pos)), dp(pos), dp(gv), dp(gc[0]), dp(gc[1]), dp(gc[2])) == 0
        exact(gv, want_v, "APIC G2P velocity")
        for a, b, nm in zip(gc, want_c, ("cX", "cY", "cZ")):
            exact(a, b, "APIC affine " + nm)
    # midpoint advection: fast particles so that the wall clamp has work on every closed / open face
    vel = np.ascontiguousarray(want_v * 5.0)
    h.set_particles(pos, vel)
    h.move_particles(dt)
    want_p, want_vm = h.get_particles()
    for var in variants(grid):
        gp, gv = pos.copy(), vel.copy()
        assert hc.hc_move_particles(C.byref(grid), dp(u), dp(v), dp(w), C.c_double(dt), C.c_double(prm.max_cfl), C.c_int(flag),
                                    C.c_int(var), C.c_uint64(len(gp)), dp(gp), dp(gv)) == 0
        exact(gp, want_p, "moved positions (variant %d)" % var); exact(gv, want_vm, "velocities after the wall clamp (variant %d)" % var)
    h.close()


@pytest.mark.parametrize("seed", range(8))
def test_collider_boundary_condition_random_grid(hc, port, seed):
    rng = np.random.default_rng(4000 + seed)
    grid = random_grid(rng)
    prm = HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=1), closed_domain_boundary_flag=int(rng.choice([63, 0b101011])),
                            max_cfl=float(rng.choice([5.0, 2.0])))
    h = port.hybrid(prm)
    lo = np.array([grid.ox, grid.oy, grid.oz]); size = np.array([grid.nx * grid.hx, grid.ny * grid.hy, grid.nz * grid.hz])
    c = lo + rng.uniform(0.3, 0.7, 3) * size
    kind = ["box", "sphere", "cylinder", "plane"][seed % 4]
    if kind == "box":
        shape = dict(kind=kind, a=tuple(c - 0.2 * size), b=tuple(c + 0.15 * size))
    elif kind == "sphere":
        shape = dict(kind=kind, a=tuple(c), b=(0.25 * size.min(), 0.0, 0.0))
    elif kind == "cylinder":
        shape = dict(kind=kind, a=tuple(c), b=(0.2 * size.min(), 0.6 * size[1], 0.0))
    else:
        shape = dict(kind=kind, a=tuple(c), b=(0.6, 0.8, 0.0))
    lin, ang, fr = tuple(rng.uniform(-0.3, 0.3, 3)), tuple(rng.uniform(-0.2, 0.2, 3)), float(rng.choice([0.0, 0.4]))
    h.set_collider_shape(**shape, linear_velocity=lin, angular_velocity=ang, friction=fr)
    nx, ny, nz = grid.res
    u = rng.uniform(-1, 1, (nz, ny, nx + 1)); v = rng.uniform(-1, 1, (nz, ny + 1, nx)); w = rng.uniform(-1, 1, (nz + 1, ny, nx))
    h.set_velocity(u, v, w)
    h.apply_boundary_condition()
    want = h.get_velocity()
    want_sdf = h.get_collider_sdf()
    arr = (ColliderShape * 1)()
    arr[0].kind = {"box": 1, "sphere": 2, "cylinder": 3, "plane": 4}[kind]
    arr[0].a = (C.c_double * 3)(*shape["a"]); arr[0].b = (C.c_double * 3)(*shape["b"])
    args = [arr, C.c_uint32(1), C.c_int(0), (C.c_double * 3)(*lin), (C.c_double * 3)(*ang), C.c_double(fr)]
    got_sdf = np.zeros_like(want_sdf)
    assert hc.hc_collider_sdf(C.byref(grid), *args, dp(got_sdf)) == 0
    exact(got_sdf, want_sdf, "collider sdf")
    gu, gv, gw = u.copy(), v.copy(), w.copy()
    assert hc.hc_apply_boundary_condition(C.byref(grid), *args, C.c_double(prm.max_cfl), C.c_int(prm.closed_domain_boundary_flag),
                                          dp(gu), dp(gv), dp(gw)) == 0
    for a, b, what in zip((gu, gv, gw), want, "uvw"):
        exact(a, b, "constrained " + what)
    h.close()


@pytest.mark.parametrize("seed", range(6))
def test_smoke_stages_random_grid(hc, port, seed):
    """Buoyancy and cubic semi-Lagrangian advection (csrc/smoke_ops.cuh) on random fields over a random grid."""
    from oracle_api import SmokeParams
    rng = np.random.default_rng(5000 + seed)
    grid = random_grid(rng)
    gravity = [(0.0, -9.8, 0.0), (3.0, -4.0, 1.0), (0.0, 0.0, -5.0)][seed % 3]
    prm = SmokeParams.make(grid, mgpcg=MgpcgParams.make(max_iterations=30, max_levels=1), gravity=gravity,
                           advection="linear" if seed % 2 else "cubic")
    s = port.smoke(prm)
    nx, ny, nz = grid.res
    den = rng.uniform(0.0, 1.0, (nz, ny, nx)); temp = rng.uniform(0.0, 5.0, (nz, ny, nx))
    hmin = min(grid.hx, grid.hy, grid.hz)
    u = rng.uniform(-1, 1, (nz, ny, nx + 1)) * 3 * hmin * 30; v = rng.uniform(-1, 1, (nz, ny + 1, nx)) * 3 * hmin * 30
    w = rng.uniform(-1, 1, (nz + 1, ny, nx)) * 3 * hmin * 30       # up to ~3 cells per step: several back-trace sub-steps
    s.set_fields(den, temp, u, v, w)
    dt = 1.0 / 30.0
    f = s.get_fields()
    s.compute_external_forces(dt)
    want = s.get_fields()
    t_sum = 0.0
    for x in f["temperature"].ravel():
        t_sum += float(x)
    gu, gv, gw = f["u"].copy(), f["v"].copy(), f["w"].copy()
    assert hc.hc_smoke_buoyancy(C.byref(prm.grid), dp(f["density"]), dp(f["temperature"]), C.c_double(prm.buoyancy_density_factor),
                                C.c_double(prm.buoyancy_temperature_factor), C.c_double(t_sum), (C.c_double * 3)(*gravity),
                                C.c_double(dt), C.c_int(prm.closed_domain_boundary_flag), dp(gu), dp(gv), dp(gw)) == 0
    for a, k in zip((gu, gv, gw), "uvw"):
        exact(a, want[k], "buoyancy " + k)
    s.compute_advection(dt)
    adv = s.get_fields()
    g = {k: want[k].copy() for k in ("density", "temperature", "u", "v", "w")}
    assert hc.hc_smoke_advect(C.byref(prm.grid), C.c_double(dt), C.c_int(prm.closed_domain_boundary_flag), C.c_int(prm.advection), dp(g["density"]),
                              dp(g["temperature"]), dp(g["u"]), dp(g["v"]), dp(g["w"])) == 0
    for k in g:
        exact(g[k], adv[k], "advected " + k)
    s.close()
