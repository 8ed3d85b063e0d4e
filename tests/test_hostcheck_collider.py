"""The CUDA source of the collider path, executed on the CPU (tests/hostcheck: the __host__ __device__ kernel bodies of
csrc/collider.cuh / collider_ops.cuh behind plain loops), against the C oracle — bit for bit.  This is what can be
verified about those kernels without a GPU: every arithmetic expression and index computation of the per-cell /
per-face / per-particle bodies; only the launch geometry is left to tests/test_gpu_collider.py."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from oracle_api import FLIP, PIC, Grid, HybridParams, MgpcgParams, have_port, load_port
from oracle_api import ColliderShape
from test_oracle_collider import ALL, COLLIDERS, EMITTERS, SETS, install

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


KINDS = {"box": 1, "sphere": 2, "cylinder": 3, "plane": 4}


def shape_args(name):
    """(shapes*, n, as_set, lin, ang, friction) of hostcheck's collider entry points for a COLLIDERS / SETS entry."""
    if name in COLLIDERS:
        c = COLLIDERS[name]
        shapes, kw, as_set = [c], c, 0
    else:
        shapes, kw = SETS[name]
        as_set = 1
    arr = (ColliderShape * len(shapes))()
    for i, sh in enumerate(shapes):
        arr[i].kind = KINDS[sh["kind"]]
        arr[i].flipped = 1 if sh.get("flipped", False) else 0
        arr[i].a = (C.c_double * 3)(*sh["a"])
        b = tuple(sh.get("b", (0.0, 0.0, 0.0))) + (0.0, 0.0, 0.0)
        arr[i].b = (C.c_double * 3)(*b[:3])
    v3 = lambda k: (C.c_double * 3)(*kw.get(k, (0.0, 0.0, 0.0)))
    return [arr, C.c_uint32(len(shapes)), C.c_int(as_set), v3("linear_velocity"), v3("angular_velocity"),
            C.c_double(kw.get("friction", 0.0))]


def dp(a):
    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
    return a.ctypes.data_as(PD)


def exact(a, b, what):
    assert np.array_equal(a, b), "%s: max abs diff %g" % (what, float(np.abs(a - b).max()))


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("n,domain", [(12, 1.0), (10, 0.9)])   # power-of-two-free spacings: the division path of jc_norm
def test_collider_sdf(hc, port, name, n, domain):
    grid = Grid.cube(n, domain)
    h = port.hybrid(HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2)))
    install(h, name)
    want = h.get_collider_sdf()
    got = np.zeros_like(want)
    assert hc.hc_collider_sdf(C.byref(grid), *shape_args(name), dp(got)) == 0
    exact(got, want, "collider sdf")
    h.close()


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("n,domain", [(12, 1.0), (16, 1.0), (10, 0.9)])   # 16: power-of-two spacing -> exact-reciprocal path
def test_apply_boundary_condition(hc, port, name, n, domain):
    grid = Grid.cube(n, domain)
    prm = HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    h = port.hybrid(prm)
    install(h, name)
    rng = np.random.default_rng(n)
    u = rng.uniform(-1, 1, (n, n, n + 1)); v = rng.uniform(-1, 1, (n, n + 1, n)); w = rng.uniform(-1, 1, (n + 1, n, n))
    h.set_velocity(u, v, w)
    h.apply_boundary_condition()
    want = h.get_velocity()
    gu, gv, gw = u.copy(), v.copy(), w.copy()
    assert hc.hc_apply_boundary_condition(C.byref(grid), *shape_args(name), C.c_double(prm.max_cfl),
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
    hc.hc_apply_boundary_condition(C.byref(grid), None, C.c_uint32(0), C.c_int(0), z, z, C.c_double(0.0), C.c_double(5.0),
                                   C.c_int(0b101011), dp(gu), dp(gv), dp(gw))
    for a, b, what in zip((gu, gv, gw), h.get_velocity(), "uvw"):
        exact(a, b, "walls " + what)
    h.close()


@pytest.mark.parametrize("name", ALL)
def test_resolve_collision(hc, port, name):
    """Collider3::resolveCollision through PicSolver3::moveParticles with dt = 0 (positions stay, collisions resolve)."""
    n = 8
    grid = Grid.cube(n)
    h = port.hybrid(HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2)))
    install(h, name)
    rng = np.random.default_rng(3)
    pos = rng.uniform(0.02, 0.98, (20000, 3))
    # plenty of particles exactly on faces / edges / axes / centres of the members
    members = [COLLIDERS[name]] if name in COLLIDERS else SETS[name][0]
    at = 0
    for c in members:
        a = np.array(c["a"], dtype=np.float64)
        if c["kind"] == "box":
            lo, hi = a, np.array(c["b"])
            pos[at:at + 200] = np.clip(pos[at:at + 200], lo, hi)
            pos[at + 200:at + 400, 0] = hi[0]; pos[at + 400:at + 600, 1] = lo[1]; pos[at + 600] = (lo + hi) / 2
        elif c["kind"] == "plane":
            nrm = np.array(c["b"])
            q = pos[at:at + 300]
            pos[at:at + 300] = q - np.outer((q - a) @ nrm, nrm)             # on the plane (up to rounding)
            pos[at + 300:at + 600] -= 0.05 * nrm                             # and some well behind it
        else:   # sphere / cylinder: the centre itself, a cloud around it, points on the axis
            pos[at] = a
            pos[at + 1:at + 300] = a + rng.normal(size=(299, 3)) * 0.05
            pos[at + 300:at + 400, 0] = a[0]; pos[at + 300:at + 400, 2] = a[2]
        at += 700
    pos = np.ascontiguousarray(np.clip(pos, 0.02, 0.98))
    vel = rng.uniform(-2, 2, pos.shape)
    h.set_particles(pos, vel)
    h.move_particles(0.0)
    want_p, want_v = h.get_particles()
    gp, gv = pos.copy(), vel.copy()
    assert hc.hc_resolve_collision(*shape_args(name), C.c_uint64(len(gp)), dp(gp), dp(gv)) == 0
    moved = int((np.abs(want_p - pos).max(axis=1) > 0).sum())
    assert moved > 100, moved
    exact(gp, want_p, "resolved positions"); exact(gv, want_v, "resolved velocities")
    h.close()


@pytest.mark.parametrize("name", ["box", "sphere", "cylinder", "tank", "cylinders", "mixed"])
def test_extrapolate_sdf_into_collider(hc, port, name):
    """The tail of buildSignedDistanceField: raw particle SDF (solver without collider) -> extrapolated (with collider)."""
    n = 12
    grid = Grid.cube(n)
    prm = HybridParams.make(grid, mode=FLIP, mgpcg=MgpcgParams.make(max_levels=2))
    raw_h, col_h = port.hybrid(prm), port.hybrid(prm)
    install(col_h, name)
    for h in (raw_h, col_h):
        h.emit_box((0.0, 0.0, 0.0), (0.7, 0.7, 1.0), 0.5 / n)
        h.build_signed_distance_field()
    raw, want = raw_h.get_sdf(), col_h.get_sdf()
    assert not np.array_equal(raw, want)
    got = raw.copy()
    assert hc.hc_extrapolate_into_collider(C.byref(grid), *shape_args(name), C.c_double(prm.max_cfl), dp(got)) == 0
    exact(got, want, "sdf extrapolated into the collider")
    raw_h.close(); col_h.close()


@pytest.mark.parametrize("lattice", ["grid", "bcc"])      # GridPointGenerator3 / BccLatticePointGenerator
@pytest.mark.parametrize("name", sorted(EMITTERS))
@pytest.mark.parametrize("n,domain", [(16, 1.0), (10, 0.9)])
def test_emit_shapes(hc, port, name, n, domain, lattice):
    """Device emission (region, lattice, acceptance) == VolumeParticleEmitter3 as restated by the oracle, point for point."""
    grid = Grid.cube(n, domain)
    h = port.hybrid(HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2)))
    shapes = EMITTERS[name]
    spacing = 0.5 * domain / n
    count = h.emit_shapes(shapes, spacing, lattice=lattice)
    bcc = C.c_int(int(lattice == "bcc"))
    want, _ = h.get_particles()
    assert count == len(want) and count > 100
    arr = (ColliderShape * len(shapes))()
    for i, sh in enumerate(shapes):
        arr[i].kind = KINDS[sh["kind"]]; arr[i].flipped = 1 if sh.get("flipped", False) else 0
        arr[i].a = (C.c_double * 3)(*sh["a"])
        b = tuple(sh.get("b", (0.0, 0.0, 0.0))) + (0.0, 0.0, 0.0)
        arr[i].b = (C.c_double * 3)(*b[:3])
    got = np.zeros_like(want)
    n_out = C.c_uint64(0)
    assert hc.hc_emit_shapes(C.byref(grid), arr, C.c_uint32(len(shapes)), C.c_int(int(len(shapes) != 1)), bcc, C.c_double(spacing), None,
                             C.byref(n_out)) == 0
    assert n_out.value == count
    assert hc.hc_emit_shapes(C.byref(grid), arr, C.c_uint32(len(shapes)), C.c_int(int(len(shapes) != 1)), bcc, C.c_double(spacing),
                             dp(got), C.byref(n_out)) == 0
    exact(got, want, "emitted positions")
    h.close()
