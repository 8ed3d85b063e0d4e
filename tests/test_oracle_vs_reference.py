"""Pins the C restatement (oracle/libjetport.so) against the UNMODIFIED reference
(oracle/_ref/libjetref.so, compiled from /root/reference by oracle/Makefile): every function of the
port must be BIT-IDENTICAL to the reference on the same inputs (the port replays the same IEEE
expressions; both are built without FMA contraction).  Skipped when the reference build is absent."""
import numpy as np
import pytest

import scenes
from oracle_api import (APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams, have_port, have_ref, load_port, load_ref)

pytestmark = pytest.mark.skipif(not (have_port() and have_ref()), reason="needs both oracle libraries")


@pytest.fixture(scope="module")
def libs():
    ref, port = load_ref(), load_port()
    ref.set_num_threads(0)
    port.set_num_threads(0)
    return ref, port


def same(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a, b), "max abs diff %g" % float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


@pytest.mark.parametrize("shape", [(12, 10, 10), (7, 5, 5), (1, 1, 1), (2, 1, 1), (16, 16, 16)])
@pytest.mark.parametrize("spd", [True, False])
def test_blas_and_relax_bit_exact(libs, shape, spd):
    ref, port = libs
    A, x, b = scenes.random_system(*shape, seed=sum(shape), spd=spd)
    same(ref.mvm(A, x), port.mvm(A, x))
    same(ref.residual(A, x, b), port.residual(A, x, b))
    same(ref.relax_red_black(A, b, x, 1.5, 3), port.relax_red_black(A, b, x, 1.5, 3))
    same(ref.relax_red_black(A, b, x, 1.0, 1), port.relax_red_black(A, b, x, 1.0, 1))
    same(ref.relax(A, b, x, 1.5, 3), port.relax(A, b, x, 1.5, 3))   # the natural ordering, FdmGaussSeidelSolver3::relax
    same(ref.relax(A, b, x, 1.0, 1), port.relax(A, b, x, 1.0, 1))
    assert ref.dot(x, b) == port.dot(x, b)
    same(ref.axpy(0.3, x, b), port.axpy(0.3, x, b))


@pytest.mark.parametrize("fine", [(16, 12, 12), (2, 2, 2), (4, 4, 4), (8, 6, 6)])
def test_restrict_correct_bit_exact(libs, fine):
    ref, port = libs
    nx, ny, nz = fine
    rng = np.random.default_rng(nx)
    f = rng.uniform(-1, 1, (nz, ny, nx))
    c = rng.uniform(-1, 1, (nz // 2, ny // 2, nx // 2))
    same(ref.restrict(f), port.restrict(f))
    same(ref.correct(c, f), port.correct(c, f))


def test_mg_num_levels(libs):
    ref, port = libs
    for dims in [(256, 256, 256), (64, 64, 64), (48, 48, 48), (10, 10, 10), (3, 3, 3), (1, 1, 1)]:
        for ml in (1, 2, 5, 6, 7, 9):
            assert ref.mg_num_levels(*dims, ml) == port.mg_num_levels(*dims, ml)


@pytest.mark.parametrize("red_black", [True, False])   # False = the natural ordering, the reference constructors' default
@pytest.mark.parametrize("n,levels", [(16, 3), (32, 4)])
def test_vcycle_and_mgpcg_bit_exact(libs, n, levels, red_black):
    ref, port = libs
    Al = scenes.poisson_levels(n, levels)
    rng = np.random.default_rng(3)
    b = rng.uniform(-1, 1, (n, n, n))
    b -= b.mean()
    x0 = rng.uniform(-1, 1, (n, n, n))
    p = MgpcgParams.make(max_iterations=25, max_levels=levels, tolerance=1e-9, use_red_black=red_black)
    same(ref.vcycle(Al, b, x0, p), port.vcycle(Al, b, x0, p))
    xr, rr = ref.mgpcg_solve(Al, b, p)
    xp, rp = port.mgpcg_solve(Al, b, p)
    same(xr, xp)
    assert rr.iterations == rp.iterations and rr.last_residual == rp.last_residual and rr.returned == rp.returned


@pytest.mark.parametrize("n,kind,collider,levels,red_black", [(16, "slab", False, 3, True), (16, "column", True, 3, True),
                                                               (32, "column", True, 5, True), (32, "slab", True, 4, True),
                                                               (24, "column", False, 9, True), (16, "column", True, 3, False),
                                                               (32, "slab", True, 4, False)])
def test_pressure_solve_bit_exact(libs, n, kind, collider, levels, red_black):
    ref, port = libs
    sc = scenes.projection_scene(n, kind, collider)
    p = MgpcgParams.make(max_iterations=200, max_levels=levels, use_red_black=red_black)
    out = []
    for o in (ref, port):
        r = o.pressure_solve(sc["grid"], sc["u"], sc["v"], sc["w"], p, collider_sdf=sc["collider_sdf"],
                             fluid_sdf=sc["fluid_sdf"])
        nl = o.mg_num_levels(n, n, n, levels)
        out.append((r, [o.pressure_get_system(sc["grid"], l) for l in range(nl)]))
    (r0, s0), (r1, s1) = out
    for a, b in zip(r0[:4], r1[:4]):
        same(a, b)
    assert r0[4].iterations == r1[4].iterations and r0[4].last_residual == r1[4].last_residual
    assert len(s0) == len(s1)
    for l in range(len(s0)):
        for key in ("fluid_sdf", "u_weights", "v_weights", "w_weights"):
            same(s0[l][key], s1[l][key])
        for fld in ("center", "right", "up", "front"):
            same(s0[l]["A"][fld], s1[l]["A"][fld])
    same(s0[0]["b"], s1[0]["b"])


@pytest.mark.parametrize("shape,spd", [((12, 10, 7), True), ((9, 9, 9), False), ((1, 1, 1), True), ((16, 16, 16), True)])
def test_iccg_bit_exact(libs, shape, spd):
    """FdmIccgSolver3 (src/jet/fdm_iccg_solver3.cpp:14-70, 152-188), Jet's default linear solver: the restatement equals the
    unmodified reference bit for bit — incomplete-Cholesky diagonal, both substitutions, solution, iteration count, residual."""
    ref, port = libs
    nx, ny, nz = shape
    A, _, b = scenes.random_system(nx, ny, nz, seed=100 + sum(shape), spd=spd)
    for x, y in zip(ref.iccg_precondition(A, b), port.iccg_precondition(A, b)):
        same(x, y)
    if spd:
        (xr, rr), (xp, rp) = ref.iccg_solve(A, b, 60, 1e-10), port.iccg_solve(A, b, 60, 1e-10)
        same(xr, xp)
        assert rr.iterations == rp.iterations and rr.last_residual == rp.last_residual and rr.returned == rp.returned


@pytest.mark.parametrize("n,kind,collider", [(16, "column", True), (24, "slab", False), (20, "column", False)])
def test_pressure_solve_iccg_and_compressed(libs, n, kind, collider):
    """GridFractionalSinglePhasePressureSolver3 with its default FdmIccgSolver3: uncompressed, port == reference bit for bit;
    with useCompressedLinearSystem(true) (what hybrid_liquid_sim sets, main.cpp:219) the reference solves the liquid cells only
    (buildCompressedSystem, …pressure_solver3.cpp:269-444) — the same system without its uncoupled identity rows — and the
    uncompressed restatement agrees with it to rounding: pressure 1e-9 relative (north_star's bound: 1e-6), iterations equal."""
    ref, port = libs
    sc = scenes.projection_scene(n, kind, collider)
    args = (sc["grid"], sc["u"], sc["v"], sc["w"])
    kw = dict(collider_sdf=sc["collider_sdf"], fluid_sdf=sc["fluid_sdf"])
    r0 = ref.pressure_solve(*args, MgpcgParams.iccg(300, 1e-9), **kw)
    p0 = port.pressure_solve(*args, MgpcgParams.iccg(300, 1e-9), **kw)
    for a, b in zip(r0[:4], p0[:4]):
        same(a, b)
    assert r0[4].iterations == p0[4].iterations and r0[4].last_residual == p0[4].last_residual
    assert r0[4].true_residual_l2 <= 1e-7 * r0[4].b_l2          # the case really converges
    rc = ref.pressure_solve(*args, MgpcgParams.iccg(300, 1e-9, compressed=True), **kw)
    pc = port.pressure_solve(*args, MgpcgParams.iccg(300, 1e-9, compressed=True), **kw)
    scale = np.abs(rc[3]).max()
    assert np.abs(rc[3] - pc[3]).max() <= 1e-9 * scale
    vscale = max(float(np.abs(a).max()) for a in args[1:])     # the two solves stop at 1e-9: velocities agree to that level
    for a, b in zip(rc[:3], pc[:3]):
        assert np.abs(a - b).max() <= 1e-7 * vscale
    assert abs(int(rc[4].iterations) - int(pc[4].iterations)) <= 1


def test_pressure_defaults_and_moving_boundary(libs):
    """NULL fluid SDF (all liquid), NULL collider, and a non-zero boundary velocity."""
    ref, port = libs
    n = 16
    sc = scenes.projection_scene(n, "slab", True)
    g = sc["grid"]
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, sc["u"].shape); v = rng.uniform(-1, 1, sc["v"].shape); w = rng.uniform(-1, 1, sc["w"].shape)
    bnd = (np.full(sc["u"].shape, 0.25), np.full(sc["v"].shape, -0.5), np.full(sc["w"].shape, 0.125))
    p = MgpcgParams.make(max_iterations=20, max_levels=3)
    for kw in (dict(), dict(collider_sdf=sc["collider_sdf"], bnd=bnd, fluid_sdf=sc["fluid_sdf"]),
               dict(collider_sdf=sc["collider_sdf"], bnd=bnd)):
        r0 = ref.pressure_solve(g, u, v, w, p, **kw)
        r1 = port.pressure_solve(g, u, v, w, p, **kw)
        for a, b in zip(r0[:4], r1[:4]):
            same(a, b)
        same(ref.pressure_get_system(g, 0)["b"], port.pressure_get_system(g, 0)["b"])


def test_bin_points(libs):
    ref, port = libs
    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.3, 1.3, (5000, 3))
    pts[:10] = 0.0
    pts[10:20] = np.array([0.11, 0.22, 0.33])  # exact multiples of the spacing
    for res, spacing in (((8, 8, 8), 0.11), ((64, 64, 64), 0.013), ((4, 2, 3), 0.5)):
        r = ref.bin_points(pts, res, spacing)
        p = port.bin_points(pts, res, spacing)
        for k in ("keys_unsorted", "keys_sorted", "start_table", "end_table"):
            same(r[k], p[k])
        # the reference's parallelSort is unstable: compare bucket membership as sets ...
        assert np.array_equal(np.sort(r["sorted_indices"]), np.sort(p["sorted_indices"]))
        assert np.array_equal(p["keys_unsorted"][p["sorted_indices"]], p["keys_sorted"])
        assert np.array_equal(r["keys_unsorted"][r["sorted_indices"]], r["keys_sorted"])
        # ... and the port's stable order against the serial searcher's buckets (ascending index)
        bucket, rank = ref.bin_points_serial(pts, res, spacing)
        same(bucket, p["keys_unsorted"])
        si, ks, st = p["sorted_indices"], p["keys_sorted"], p["start_table"]
        assert np.array_equal(rank[si], np.arange(len(si), dtype=np.uint64) - st[ks])


def test_bin_points_empty(libs):
    ref, port = libs
    pts = np.zeros((0, 3))
    r = ref.bin_points(pts, (4, 4, 4), 0.1)
    p = port.bin_points(pts, (4, 4, 4), 0.1)
    same(r["start_table"], p["start_table"])
    same(r["end_table"], p["end_table"])


@pytest.mark.parametrize("mode", [PIC, FLIP, APIC])
def test_hybrid_stages_bit_exact(libs, mode):
    """Stage by stage, 3 sub-steps of a 16^3 dam break: the port's serial P2G and the reference's are the same
    summation order, so even the transferred grids are bit-identical."""
    ref, port = libs
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=3),
                            pic_blending_factor=(0.1 if mode == FLIP else 0.0))
    hr, hp = ref.hybrid(prm), port.hybrid(prm)
    assert hr.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n) == hp.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n)
    same(hr.get_particles()[0], hp.get_particles()[0])
    dt = 1 / 60.0
    for _ in range(3):
        assert hr.cfl(dt) == hp.cfl(dt)
        assert hr.num_sub_steps(dt) == hp.num_sub_steps(dt)
        hr.begin_advance_time_step(dt); hp.begin_advance_time_step(dt)
        for a, b in zip(hr.get_velocity(), hp.get_velocity()):
            same(a, b)
        same(hr.get_sdf(), hp.get_sdf())
        for a, b in zip(hr.get_markers(), hp.get_markers()):
            same(a, b)
        hr.compute_external_forces(dt); hp.compute_external_forces(dt)
        for a, b in zip(hr.get_velocity(), hp.get_velocity()):
            same(a, b)
        r0, r1 = hr.compute_pressure(dt), hp.compute_pressure(dt)
        assert r0.iterations == r1.iterations and r0.last_residual == r1.last_residual
        for a, b in zip(hr.get_velocity(), hp.get_velocity()):
            same(a, b)
        same(hr.get_pressure(), hp.get_pressure())
        hr.compute_advection(dt); hp.compute_advection(dt)
        for a, b in zip(hr.get_particles(), hp.get_particles()):
            same(a, b)
        if mode == APIC:
            for a, b in zip(hr.get_affine(), hp.get_affine()):
                same(a, b)


def test_hybrid_frames_bit_exact(libs):
    """Two whole frames through the adaptive sub-stepping loop (onAdvanceTimeStep as a unit)."""
    ref, port = libs
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=3))
    hr, hp = ref.hybrid(prm), port.hybrid(prm)
    pts = scenes.dam_particles(n)
    hr.set_particles(pts); hp.set_particles(pts)
    for _ in range(2):
        n0, s0 = hr.advance_frame(1 / 30.0)
        n1, s1 = hp.advance_frame(1 / 30.0)
        assert n0 == n1 and s0.pcg_iterations == s1.pcg_iterations
        for a, b in zip(hr.get_particles(), hp.get_particles()):
            same(a, b)
