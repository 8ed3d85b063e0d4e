"""The CUDA source of the collider path, executed on the CPU (tests/hostcheck: the __host__ __device__ kernel bodies of
csrc/collider.cuh / collider_ops.cuh behind plain loops), against the C oracle — bit for bit.  This is what can be
verified about those kernels without a GPU: every arithmetic expression and index computation of the per-cell /
per-face / per-particle bodies; only the launch geometry is left to tests/firstrun/zz_collider.py."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from oracle_api import FLIP, PIC, Grid, HybridParams, MgpcgParams, have_port, load_port
from test_oracle_collider import COLLIDERS

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_hostcheck():
    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(HERE, "hostcheck", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = mod.build()
    return C.CDLL(lib) if lib else None


pytestmark = pytest.mark.skipif(not have_port(), reason="oracle/libjetport.so not built")
PD = C.POINTER(C.c_double)


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


def shape_args(c):
    kind = {"box": 1, "sphere": 2}[c["kind"]]
    v3 = lambda k: (C.c_double * 3)(*c.get(k, (0.0, 0.0, 0.0)))
    return [C.c_int(kind), v3("a"), v3("b"), v3("linear_velocity"), v3("angular_velocity"), C.c_double(c.get("friction", 0.0))]


def dp(a):
    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
    return a.ctypes.data_as(PD)


def exact(a, b, what):
    assert np.array_equal(a, b), "%s: max abs diff %g" % (what, float(np.abs(a - b).max()))


@pytest.mark.parametrize("name", sorted(COLLIDERS))
@pytest.mark.parametrize("n,domain", [(12, 1.0), (10, 0.9)])   # power-of-two-free spacings: the division path of jc_norm
def test_collider_sdf(hc, port, name, n, domain):
    grid = Grid.cube(n, domain)
    h = port.hybrid(HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2)))
    h.set_collider_shape(**COLLIDERS[name])
    want = h.get_collider_sdf()
    got = np.zeros_like(want)
    assert hc.hc_collider_sdf(C.byref(grid), *shape_args(COLLIDERS[name]), dp(got)) == 0
    exact(got, want, "collider sdf")
    h.close()


@pytest.mark.parametrize("name", sorted(COLLIDERS))
@pytest.mark.parametrize("n,domain", [(12, 1.0), (16, 1.0), (10, 0.9)])   # 16: power-of-two spacing -> exact-reciprocal path
def test_apply_boundary_condition(hc, port, name, n, domain):
    grid = Grid.cube(n, domain)
    prm = HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    h = port.hybrid(prm)
    h.set_collider_shape(**COLLIDERS[name])
    rng = np.random.default_rng(n)
    u = rng.uniform(-1, 1, (n, n, n + 1)); v = rng.uniform(-1, 1, (n, n + 1, n)); w = rng.uniform(-1, 1, (n + 1, n, n))
    h.set_velocity(u, v, w)
    h.apply_boundary_condition()
    want = h.get_velocity()
    gu, gv, gw = u.copy(), v.copy(), w.copy()
    assert hc.hc_apply_boundary_condition(C.byref(grid), *shape_args(COLLIDERS[name]), C.c_double(prm.max_cfl),
                                          C.c_int(prm.closed_domain_boundary_flag), dp(gu), dp(gv), dp(gw)) == 0
    for a, b, what in zip((gu, gv, gw), want, "uvw"):
        exact(a, b, "constrained " + what)
    h.close()


def test_walls_only(hc, port):
    n = 8
    grid = Grid.cube(n)
    prm = HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2), closed_domain_boundary_flag=0b101011)
    h = port.hybrid(prm)
    rng = np.random.default_rng(1)
    u = rng.uniform(-1, 1, (n, n, n + 1)); v = rng.uniform(-1, 1, (n, n + 1, n)); w = rng.uniform(-1, 1, (n + 1, n, n))
    h.set_velocity(u, v, w)
    h.apply_boundary_condition()
    z = (C.c_double * 3)(0, 0, 0)
    gu, gv, gw = u.copy(), v.copy(), w.copy()
    hc.hc_apply_boundary_condition(C.byref(grid), C.c_int(0), z, z, z, z, C.c_double(0.0), C.c_double(5.0),
                                   C.c_int(0b101011), dp(gu), dp(gv), dp(gw))
    for a, b, what in zip((gu, gv, gw), h.get_velocity(), "uvw"):
        exact(a, b, "walls " + what)
    h.close()


@pytest.mark.parametrize("name", ["box", "box_moving", "sphere", "box_offgrid"])
def test_resolve_collision(hc, port, name):
    """Collider3::resolveCollision through PicSolver3::moveParticles with dt = 0 (positions stay, collisions resolve)."""
    n = 8
    grid = Grid.cube(n)
    h = port.hybrid(HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2)))
    h.set_collider_shape(**COLLIDERS[name])
    rng = np.random.default_rng(3)
    pos = rng.uniform(0.02, 0.98, (20000, 3))
    # plenty of particles exactly on faces / edges / the centre of the shape
    c = COLLIDERS[name]
    if c["kind"] == "box":
        lo, hi = np.array(c["a"]), np.array(c["b"])
        pos[:200] = np.clip(pos[:200], lo, hi)
        pos[200:400, 0] = hi[0]; pos[400:600, 1] = np.clip(lo[1], 0.02, 0.98); pos[600] = np.clip((lo + hi) / 2, 0.02, 0.98)
    else:
        pos[0] = c["a"]; pos[1:300] = np.array(c["a"]) + rng.normal(size=(299, 3)) * 0.05
    pos = np.ascontiguousarray(np.clip(pos, 0.02, 0.98))
    vel = rng.uniform(-2, 2, pos.shape)
    h.set_particles(pos, vel)
    h.move_particles(0.0)
    want_p, want_v = h.get_particles()
    gp, gv = pos.copy(), vel.copy()
    assert hc.hc_resolve_collision(*shape_args(c), C.c_uint64(len(gp)), dp(gp), dp(gv)) == 0
    moved = int((np.abs(want_p - pos).max(axis=1) > 0).sum())
    assert moved > 100, moved
    exact(gp, want_p, "resolved positions"); exact(gv, want_v, "resolved velocities")
    h.close()


@pytest.mark.parametrize("name", ["box", "sphere"])
def test_extrapolate_sdf_into_collider(hc, port, name):
    """The tail of buildSignedDistanceField: raw particle SDF (solver without collider) -> extrapolated (with collider)."""
    n = 12
    grid = Grid.cube(n)
    prm = HybridParams.make(grid, mode=FLIP, mgpcg=MgpcgParams.make(max_levels=2))
    raw_h, col_h = port.hybrid(prm), port.hybrid(prm)
    col_h.set_collider_shape(**COLLIDERS[name])
    for h in (raw_h, col_h):
        h.emit_box((0.0, 0.0, 0.0), (0.7, 0.7, 1.0), 0.5 / n)
        h.build_signed_distance_field()
    raw, want = raw_h.get_sdf(), col_h.get_sdf()
    assert not np.array_equal(raw, want)
    got = raw.copy()
    assert hc.hc_extrapolate_into_collider(C.byref(grid), *shape_args(COLLIDERS[name]), C.c_double(prm.max_cfl), dp(got)) == 0
    exact(got, want, "sdf extrapolated into the collider")
    raw_h.close(); col_h.close()
