"""GPU parity, seam 1: FdmBlas3 / red-black SOR / MG transfer / V-cycle / MGPCG through the C ABI vs the oracle.
Everything without a reduction must be BIT-EXACT; dot products and PCG are compared within stated tolerances."""
import numpy as np
import pytest

import golden_cases as gc
import scenes
from gpu_common import checker, jetcu_or_fail
from oracle_api import MgpcgParams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jc():
    return jetcu_or_fail()


@pytest.fixture(scope="module")
def oracle():
    return checker()


def same(a, b, what=""):
    assert np.array_equal(a, b), "%s not bit-identical: max abs diff %g" % (what, float(np.abs(a - b).max()))


def test_library_identity(jc):
    assert "sm_100a" in jc.version()


def test_golden_linear_ops(jc):
    gc.linear_ops(jc, dot_rtol=1e-15)


def test_golden_mgpcg_poisson(jc):
    gc.mgpcg_poisson(jc, exact_bits=False)


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 1, 1), (7, 5, 5), (12, 10, 10), (33, 17, 17), (64, 64, 64),
                                   (130, 66, 66), (257, 9, 9)])
@pytest.mark.parametrize("spd", [True, False])
def test_stencils_bit_exact(jc, oracle, shape, spd):
    A, x, b = scenes.random_system(*shape, seed=sum(shape), spd=spd)
    same(jc.mvm(A, x), oracle.mvm(A, x), "mvm")
    same(jc.residual(A, x, b), oracle.residual(A, x, b), "residual")
    same(jc.relax_red_black(A, b, x, 1.5, 3), oracle.relax_red_black(A, b, x, 1.5, 3), "relaxRedBlack x3")
    same(jc.relax_red_black(A, b, x, 1.0, 1), oracle.relax_red_black(A, b, x, 1.0, 1), "relaxRedBlack w=1")
    # the natural (lexicographic) ordering, FdmGaussSeidelSolver3::relax: serial in the reference, a tile-hyperplane wavefront
    # of pipelined sweeps on the device (fdm.cu) — the same operands per cell, hence the same bits
    same(jc.relax(A, b, x, 1.5, 3), oracle.relax(A, b, x, 1.5, 3), "relax (natural ordering) x3")
    same(jc.relax(A, b, x, 1.0, 1), oracle.relax(A, b, x, 1.0, 1), "relax (natural ordering) w=1")
    if spd:
        same(jc.relax(A, b, x, 1.3, 35), oracle.relax(A, b, x, 1.3, 35), "relax (natural ordering) x35: two launch groups")
    same(jc.axpy(0.3, x, b), oracle.axpy(0.3, x, b), "axpy")
    d0, d1 = jc.dot(x, b), oracle.dot(x, b)
    assert abs(d0 - d1) <= 1e-14 * np.abs(x * b).sum() + 1e-300


@pytest.mark.parametrize("fine", [(2, 2, 2), (4, 4, 4), (16, 12, 12), (8, 6, 6), (64, 64, 64), (130, 34, 34)])
def test_restrict_correct_bit_exact(jc, oracle, fine):
    nx, ny, nz = fine
    rng = np.random.default_rng(nx)
    f = rng.uniform(-1, 1, (nz, ny, nx))
    c = rng.uniform(-1, 1, (nz // 2, ny // 2, nx // 2))
    same(jc.restrict(f), oracle.restrict(f), "restrict")
    same(jc.correct(c, f), oracle.correct(c, f), "correct")


def test_dot_deterministic(jc):
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, 1 << 20)
    b = rng.uniform(-1, 1, 1 << 20)
    vals = {jc.dot(a, b) for _ in range(5)}
    assert len(vals) == 1


@pytest.mark.parametrize("red_black", [True, False])   # False: the reference constructors' default ordering
@pytest.mark.parametrize("n,levels", [(16, 3), (32, 4), (64, 5)])
def test_vcycle_bit_exact(jc, oracle, n, levels, red_black):
    Al = scenes.poisson_levels(n, levels)
    rng = np.random.default_rng(3)
    b = rng.uniform(-1, 1, (n, n, n)); b -= b.mean()
    x0 = rng.uniform(-1, 1, (n, n, n))
    p = MgpcgParams.make(max_levels=levels, use_red_black=red_black)
    same(jc.vcycle(Al, b, x0, p), oracle.vcycle(Al, b, x0, p), "V-cycle")


@pytest.mark.parametrize("red_black", [True, False])
@pytest.mark.parametrize("n,levels,shift", [(32, 4, 0.05), (32, 3, 0.5), (64, 5, 0.01)])
def test_mgpcg_vs_oracle(jc, oracle, n, levels, shift, red_black):
    """SPD (regularised) Poisson systems.  Where the oracle's TRUE residual is small (the reference's PCG can leave
    early on sigma <= tol^2) x must agree within 1e-6 relative L-inf; in every case the iteration
    count and the exit sigma must follow the oracle (same control flow)."""
    Al = scenes.poisson_levels(n, levels)
    for A in Al:
        A["center"] += shift
    rng = np.random.default_rng(7)
    b = rng.uniform(-1, 1, (n, n, n))
    p = MgpcgParams.make(max_iterations=100, max_levels=levels, tolerance=1e-9, use_red_black=red_black)
    x0, r0 = jc.mgpcg_solve(Al, b, p)
    x1, r1 = oracle.mgpcg_solve(Al, b, p)
    assert abs(int(r0.iterations) - int(r1.iterations)) <= 1
    if r1.true_residual_l2 <= 1e-7 * r1.b_l2:
        assert np.abs(x0 - x1).max() <= 1e-6 * np.abs(x1).max()
        assert r0.true_residual_l2 <= 1e-7 * r0.b_l2
    else:
        assert int(r0.iterations) == int(r1.iterations)
        assert np.abs(x0 - x1).max() <= 1e-5 * np.abs(x1).max()
        assert abs(r0.true_residual_l2 - r1.true_residual_l2) <= 1e-3 * r1.true_residual_l2


@pytest.mark.parametrize("shape,spd", [((1, 1, 1), True), ((9, 3, 26), False), ((20, 17, 9), True), ((33, 17, 40), True), ((64, 64, 64), True)])
def test_iccg_preconditioner_bit_exact(jc, oracle, shape, spd):
    """FdmIccgSolver3::Preconditioner (src/jet/fdm_iccg_solver3.cpp:14-70): the reference's lexicographic recurrences, run on the
    device as tile-hyperplane sweeps (csrc/iccg.cu) — the same per-cell arithmetic, hence the same bits (no reduction)."""
    nx, ny, nz = shape
    A, _, b = scenes.random_system(nx, ny, nz, seed=40 + sum(shape), spd=spd)
    for got, want, what in zip(jc.iccg_precondition(A, b), oracle.iccg_precondition(A, b), ("IC diagonal", "forward", "backward")):
        same(got, want, what)


@pytest.mark.parametrize("n", [16, 40])
def test_iccg_vs_oracle(jc, oracle, n):
    """FdmIccgSolver3::solve (:152-188): same iteration count (dot products are a fixed tree here, serial sums in the reference:
    +-1 at the tolerance threshold), solution to 1e-9 of its scale."""
    A = scenes.poisson_levels(n, 1)[0]
    A["center"] += 0.05
    b = np.random.default_rng(n).uniform(-1, 1, (n, n, n))
    (xg, rg), (xo, ro) = jc.iccg_solve(A, b, 200, 1e-9), oracle.iccg_solve(A, b, 200, 1e-9)
    assert abs(int(rg.iterations) - int(ro.iterations)) <= 1 and rg.returned == ro.returned == 1
    assert ro.true_residual_l2 <= 1e-8 * ro.b_l2
    assert np.abs(xg - xo).max() <= 1e-9 * np.abs(xo).max()
    # the iteration cap as the exit (Jet's default is 100 iterations: large systems leave through it)
    (xg, rg), (xo, ro) = jc.iccg_solve(A, b, 5, 1e-12), oracle.iccg_solve(A, b, 5, 1e-12)
    assert rg.iterations == ro.iterations == 5 and rg.returned == ro.returned == 0
    assert np.abs(xg - xo).max() <= 1e-12 * np.abs(xo).max() and abs(rg.last_residual - ro.last_residual) <= 1e-10 * ro.last_residual


def test_errors(jc):
    from fluid_engine_dev_b200 import JetcuError
    A, x, b = scenes.random_system(4, 4, 4)
    with pytest.raises(JetcuError) as e:
        jc.restrict(np.zeros((3, 4, 4)))
    assert e.value.rc == 1  # JETCU_ERR_INVALID_ARG, like JET_ASSERT(finer == 2*coarser)
