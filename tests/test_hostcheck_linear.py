"""The linear-system and pressure-system kernels' bodies (csrc/fdm_ops.cuh, pressure_ops.cuh) executed on the CPU through
tests/hostcheck, against the C oracle — bit for bit:
  * FdmBlas3 mvm / residual, red-black SOR, MG restrict / correct on host-supplied matrices (the SoA form);
  * the pressure system build (weights, f32 restriction, 7-point matrix, rhs) in the PACKED form the solver runs on
    (diagonal + face weights + one code byte per cell), expanded back to Jet's rows on every MG level;
  * mvm / residual / SOR on the packed matrix in both smoother forms (float weights, face codes with the bulk-diagonal
    shortcut) — i.e. the arithmetic of k_relax_colour_code, the kernel a FLIP step spends a third of its time in;
  * applyPressureGradient."""
import ctypes as C

import numpy as np
import pytest

import scenes
from oracle_api import ROW_DTYPE, MgpcgParams, Row, have_port, load_port
from test_hostcheck_collider import _load_hostcheck, dp, exact

pytestmark = pytest.mark.skipif(not have_port(), reason="oracle/libjetport.so not built")
PR = C.POINTER(Row)


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


def rp(a):
    return a.ctypes.data_as(PR)


@pytest.mark.parametrize("shape", [(12, 10, 10), (7, 5, 5), (1, 1, 1), (16, 16, 16)])
@pytest.mark.parametrize("spd", [True, False])
def test_blas_and_relax(hc, port, shape, spd):
    nx, ny, nz = shape
    A, x, b = scenes.random_system(nx, ny, nz, seed=sum(shape), spd=spd)
    out = np.zeros_like(x)
    hc.hc_mvm(nx, ny, nz, rp(A), dp(x), dp(out)); exact(out, port.mvm(A, x), "mvm")
    hc.hc_residual(nx, ny, nz, rp(A), dp(x), dp(b), dp(out)); exact(out, port.residual(A, x, b), "residual")
    for sor, it in ((1.5, 3), (1.0, 1)):
        xs = x.copy()
        hc.hc_relax(nx, ny, nz, rp(A), dp(b), C.c_double(sor), C.c_uint(it), dp(xs))
        exact(xs, port.relax_red_black(A, b, x, sor, it), "relax sor=%g" % sor)


@pytest.mark.parametrize("shape", [(20, 17, 9), (8, 8, 8), (1, 1, 1), (9, 3, 26), (16, 16, 16)])
@pytest.mark.parametrize("spd", [True, False])
def test_iccg_preconditioner_tile_hyperplanes(hc, port, shape, spd):
    """FdmIccgSolver3::Preconditioner (src/jet/fdm_iccg_solver3.cpp:14-70) is three lexicographic recurrences; iccg.cu sweeps
    8^3 tiles in tile-hyperplane order and the cells of a tile in cell-hyperplane order.  The traversal, replayed on the CPU
    with the kernel's cell bodies, gives the bits of the serial loops (oracle port == unmodified reference)."""
    nx, ny, nz = shape
    A, _, b = scenes.random_system(nx, ny, nz, seed=7 + sum(shape), spd=spd)
    d, y, x = np.zeros_like(b), np.zeros_like(b), np.zeros_like(b)
    assert hc.hc_iccg_precondition(nx, ny, nz, rp(A), dp(b), dp(d), dp(y), dp(x)) == 0
    wd, wy, wx = port.iccg_precondition(A, b)
    exact(d, wd, "IC diagonal"); exact(y, wy, "forward substitution"); exact(x, wx, "backward substitution")


@pytest.mark.parametrize("shape", [(20, 17, 9), (8, 8, 8), (1, 1, 1), (9, 3, 26), (16, 16, 16), (33, 32, 24), (40, 24, 24)])
@pytest.mark.parametrize("spd", [True, False])
def test_lexicographic_relax_schedules(hc, port, shape, spd):
    """FdmGaussSeidelSolver3::relax (src/jet/fdm_gauss_seidel_solver3.cpp:103-120) is one serial loop; fdm.cu runs it as a
    wavefront: cell hyperplanes of pipelined sweeps in one CTA on small levels (<= 8192 cells), tile hyperplanes of pipelined
    sweeps (two apart) on large ones.  Replayed on the CPU with the kernel's cell body, visiting the concurrent CTAs / cells of
    a launch / step in ascending and in descending order, both schedules give the bits of the serial loop."""
    nx, ny, nz = shape
    A, x, b = scenes.random_system(nx, ny, nz, seed=11 + sum(shape), spd=spd)
    for sor, it in ((1.5, 1), (1.5, 5), (1.0, 2), (1.3, 35)):
        if it > 5 and (not spd or (nx * ny * nz > 8192 and nx * ny * nz != 33 * 32 * 24)):
            continue   # the > 32 sweeps case (two launch groups): convergent systems, one large and every small shape
        want = port.relax(A, b, x, sor, it)
        for order in (0, 1, 2, 3):   # ascending, descending, two pseudo-random permutations of a launch's CTAs
            xs = x.copy()
            assert hc.hc_relax_lex(nx, ny, nz, rp(A), dp(b), C.c_double(sor), C.c_uint(it), C.c_int(order), dp(xs)) == 0
            exact(xs, want, "lexicographic relax sor=%g sweeps=%d order=%d" % (sor, it, order))


@pytest.mark.parametrize("fine", [(16, 12, 12), (2, 2, 2), (8, 6, 6), (8, 16, 8), (12, 8, 6), (6, 4, 10)])   # the last three: ny != nz (extension)
def test_restrict_correct(hc, port, fine):
    nx, ny, nz = fine
    rng = np.random.default_rng(nx)
    f = rng.uniform(-1, 1, (nz, ny, nx)); c = rng.uniform(-1, 1, (nz // 2, ny // 2, nx // 2))
    got = np.zeros_like(c)
    hc.hc_restrict(nx, ny, nz, dp(f), dp(got)); exact(got, port.restrict(f), "restrict")
    got = f.copy()
    hc.hc_correct(nx, ny, nz, dp(c), dp(got)); exact(got, port.correct(c, f), "correct")


@pytest.mark.parametrize("n,kind,collider,domain", [(32, "slab", True, 1.0), (32, "column", True, 1.0), (24, "column", False, 0.9),
                                                    (16, "slab", True, 1.0)])
def test_packed_pressure_system(hc, port, n, kind, collider, domain):
    sc = scenes.projection_scene(n, kind, collider)
    grid = sc["grid"]
    if domain != 1.0:
        grid.hx = grid.hy = grid.hz = domain / n
    max_levels = 4
    prm = MgpcgParams.make(max_iterations=40, max_levels=max_levels)
    uo, vo, wo, p, res = port.pressure_solve(grid, sc["u"], sc["v"], sc["w"], prm, collider_sdf=sc["collider_sdf"], fluid_sdf=sc["fluid_sdf"])
    none = C.POINTER(C.c_double)()
    csdf = dp(np.ascontiguousarray(sc["collider_sdf"])) if sc["collider_sdf"] is not None else none
    levels = hc.hc_pressure_build(C.byref(grid), C.c_uint32(max_levels), dp(sc["fluid_sdf"]), csdf, dp(sc["u"]), dp(sc["v"]), dp(sc["w"]),
                                  none, none, none)
    assert levels == port.mg_num_levels(n, n, n, max_levels)
    rng = np.random.default_rng(n)
    for l in range(levels):
        want = port.pressure_get_system(grid, l)
        m = n >> l
        A = np.zeros((m, m, m), dtype=ROW_DTYPE); b = np.zeros((m, m, m))
        assert hc.hc_pressure_get_rows(C.c_uint32(l), rp(A), dp(b)) == 0
        for fld in ("center", "right", "up", "front"):
            exact(A[fld], want["A"][fld], "level %d A.%s" % (l, fld))
        if l == 0:
            exact(b, want["b"], "rhs")
        # the packed kernels against the oracle's kernels on the SAME matrix
        x = rng.uniform(-1, 1, (m, m, m)); rhs = rng.uniform(-1, 1, (m, m, m)); out = np.zeros_like(x)
        hc.hc_packed_mvm(C.c_uint32(l), dp(x), dp(out)); exact(out, port.mvm(want["A"], x), "packed mvm level %d" % l)
        hc.hc_packed_residual(C.c_uint32(l), dp(x), dp(rhs), dp(out)); exact(out, port.residual(want["A"], x, rhs), "packed residual level %d" % l)
        ref = port.relax_red_black(want["A"], rhs, x, 1.5, 3)
        for codes in (0, 1):
            xs = x.copy()
            hc.hc_packed_relax(C.c_uint32(l), dp(rhs), C.c_double(1.5), C.c_uint(3), C.c_int(codes), dp(xs))
            exact(xs, ref, "packed SOR level %d (%s)" % (l, "face codes + bulk diagonal" if codes else "float weights"))
        # the natural ordering on the packed level (k_relax_lex_tiles with face codes / k_relax_lex_small), both visiting orders
        lex = port.relax(want["A"], rhs, x, 1.5, 3)
        for order in (0, 1):
            xs = x.copy()
            hc.hc_packed_relax_lex(C.c_uint32(l), dp(rhs), C.c_double(1.5), C.c_uint(3), C.c_int(order), dp(xs))
            exact(xs, lex, "packed lexicographic relax level %d order %d" % (l, order))
        # the L2-resident chunk-interleaved sweep order (JETCU_RELAX_CHUNK) must give the bits of the two full passes
        for chunk in (1, 3, 5, m):
            xs = x.copy()
            hc.hc_packed_relax_chunked(C.c_uint32(l), dp(rhs), C.c_double(1.5), C.c_uint(3), C.c_int(chunk), dp(xs))
            exact(xs, ref, "chunk-interleaved SOR level %d chunk %d" % (l, chunk))
        # several sweeps as one skewed wavefront over z-chunks (JETCU_RELAX_SKEW): still the bits of the full passes
        ref5 = port.relax_red_black(want["A"], rhs, x, 1.5, 5)
        for chunk, skew in ((1, 2), (2, 5), (3, 3), (4, 2), (m, 5)):
            xs = x.copy()
            hc.hc_packed_relax_skewed(C.c_uint32(l), dp(rhs), C.c_double(1.5), C.c_uint(5), C.c_int(chunk), C.c_int(skew), dp(xs))
            exact(xs, ref5, "skewed SOR wavefront level %d chunk %d skew %d" % (l, chunk, skew))
        # a slab's communication-avoiding block of k sweeps (redundant halo windows shrinking by a plane per pass) as one
        # wavefront: the owned planes carry the bits of k undivided sweeps
        if m >= 8:
            for k, chunk, (z0, z1) in ((2, 2, (0, m // 2)), (2, 3, (m // 2, m)), (1, 1, (m // 4, m // 2)), (3, 2, (m // 4, 3 * m // 4))):
                refk = port.relax_red_black(want["A"], rhs, x, 1.5, k)
                xs = x.copy()
                hc.hc_packed_relax_skewed_slab(C.c_uint32(l), dp(rhs), C.c_double(1.5), C.c_int(k), C.c_int(chunk), C.c_int(z0), C.c_int(z1), dp(xs))
                exact(xs[z0:z1], refk[z0:z1], "slab block wavefront level %d k %d chunk %d planes %d..%d" % (l, k, chunk, z0, z1))
    # applyPressureGradient with the oracle's pressure
    gu, gv, gw = sc["u"].copy(), sc["v"].copy(), sc["w"].copy()
    hc.hc_apply_gradient(dp(p), dp(sc["u"]), dp(sc["v"]), dp(sc["w"]), dp(gu), dp(gv), dp(gw))
    exact(gu, uo, "projected u"); exact(gv, vo, "projected v"); exact(gw, wo, "projected w")
