"""Known-answer tests taken from the reference's own unit tests, run against the oracles."""
import numpy as np
import pytest

import scenes
from oracle_api import Grid, MgpcgParams, have_port, have_ref, load_port, load_ref, size3


def _oracles():
    out = []
    if have_port():
        out.append("port")
    if have_ref():
        out.append("reference")
    return out


@pytest.fixture(scope="module", params=_oracles())
def oracle(request):
    o = load_port() if request.param == "port" else load_ref()
    o.set_num_threads(0)
    return o


def free_surface_case():
    """GridFractionalSinglePhasePressureSolver3.SolveFreeSurface
    (src/tests/unit_tests/grid_fractional_single_phase_pressure_solver3_tests.cpp:17-74): 3^3 grid, h = 1,
    v = 1 on interior faces, fluidSdf = y - 2  ->  velocity 0, pressure = 1.5 - j for j < 2."""
    grid = Grid(3, 3, 3, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0)
    u = np.zeros(size3(grid.usize)); v = np.zeros(size3(grid.vsize)); w = np.zeros(size3(grid.wsize))
    v[:, 1:3, :] = 1.0
    _, Y, _ = scenes.cell_centers(grid)
    return grid, u, v, w, np.ascontiguousarray(Y - 2.0)


def check_free_surface(api):
    grid, u, v, w, sdf = free_surface_case()
    p = MgpcgParams.make(max_iterations=100, max_levels=1, tolerance=1e-9)
    uo, vo, wo, pr, res = api.pressure_solve(grid, u, v, w, p, fluid_sdf=sdf)
    assert np.abs(uo).max() < 1e-6 and np.abs(vo).max() < 1e-6 and np.abs(wo).max() < 1e-6
    for j in range(2):
        assert np.abs(pr[:, j, :] - (1.5 - j)).max() < 1e-6


def test_solve_free_surface(oracle):
    check_free_surface(oracle)


def test_solve_free_surface_default_iccg(oracle):
    """The reference's test as written: a default-constructed solver, i.e. FdmIccgSolver3(100, 1e-6) (…pressure_solver3.cpp:450),
    with and without the compressed system (…solver3_tests.cpp:17-74 and :76-133)."""
    grid, u, v, w, sdf = free_surface_case()
    for compressed in (False, True):
        uo, vo, wo, pr, res = oracle.pressure_solve(grid, u, v, w, MgpcgParams.iccg(100, 1e-6, compressed=compressed), fluid_sdf=sdf)
        assert np.abs(uo).max() < 1e-6 and np.abs(vo).max() < 1e-6 and np.abs(wo).max() < 1e-6
        for j in range(2):
            assert np.abs(pr[:, j, :] - (1.5 - j)).max() < 1e-6


def check_relax_decreases(api):
    """FdmGaussSeidelSolver3.RelaxRedBlack (src/tests/unit_tests/fdm_gauss_seidel_solver3_tests.cpp:64-84):
    the residual norm decreases monotonically over red-black sweeps on a Poisson system."""
    n = 16
    A = scenes.poisson_levels(n, 1)[0]
    A["center"] += 0.0
    rng = np.random.default_rng(1)
    b = rng.uniform(-1, 1, (n, n, n)); b -= b.mean()
    x = np.zeros((n, n, n))
    last = np.linalg.norm(api.residual(A, x, b))
    for _ in range(20):
        x = api.relax_red_black(A, b, x, 1.0, 5)
        cur = np.linalg.norm(api.residual(A, x, b))
        assert cur < last
        last = cur


def test_relax_red_black_decreases(oracle):
    check_relax_decreases(oracle)


def check_vcycle_reduces(api):
    """FdmMgSolver3.Solve (src/tests/unit_tests/fdm_mg_solver3_tests.cpp:13-69): one V-cycle reduces the residual."""
    n, levels = 32, 4
    Al = scenes.poisson_levels(n, levels)
    rng = np.random.default_rng(2)
    b = rng.uniform(-1, 1, (n, n, n)); b -= b.mean()
    x0 = np.zeros((n, n, n))
    p = MgpcgParams.make(max_levels=levels)
    x1 = api.vcycle(Al, b, x0, p)
    assert np.linalg.norm(api.residual(Al[0], x1, b)) < 0.05 * np.linalg.norm(b)


def test_vcycle_reduces_residual(oracle):
    check_vcycle_reduces(oracle)


def check_bucket_counts(api):
    """PointParallelHashGridSearcher3.Build (src/tests/unit_tests/point_parallel_hash_grid_searcher3_tests.cpp:59-98):
    per-bucket counts from the start/end tables equal a direct histogram of the keys."""
    pts = np.random.default_rng(4).uniform(0, 1, (4000, 3))
    res = (4, 4, 4)
    r = api.bin_points(pts, res, 0.25)
    hist = np.bincount(r["keys_unsorted"].astype(np.int64), minlength=64)
    st, en = r["start_table"], r["end_table"]
    empty = np.uint64(2 ** 64 - 1)
    for key in range(64):
        if hist[key] == 0:
            assert st[key] == empty and en[key] == empty
        else:
            assert int(en[key] - st[key]) == hist[key]
    b = np.floor(pts / 0.25).astype(np.int64)
    assert np.array_equal(r["keys_unsorted"].astype(np.int64), (b[:, 2] * 4 + b[:, 1]) * 4 + b[:, 0])


def test_bucket_counts(oracle):
    check_bucket_counts(oracle)
