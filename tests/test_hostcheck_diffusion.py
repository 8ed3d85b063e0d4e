"""The product's diffusion kernel bodies (csrc/diffusion_ops.cuh: markers, matrix rows, right-hand side of
GridBackwardEulerDiffusionSolver3) executed on the CPU through tests/hostcheck, against the C oracle — bit for bit.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle_api import ROW_DTYPE, Row, load_port
from test_hostcheck_collider import _load_hostcheck, dp
from test_oracle_diffusion import fields


@pytest.fixture(scope="module")
def hc():
    lib = _load_hostcheck()
    if lib is None:
        pytest.skip("nvcc not available to build tests/hostcheck")
    return lib


@pytest.fixture(scope="module")
def port():
    return load_port()


@pytest.mark.parametrize("boundary_type", [0, 1])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("with_collider,with_fluid", [(True, True), (False, True), (True, False)])
def test_system_and_solution(hc, port, kind, boundary_type, with_collider, with_fluid):
    grid, u, v, w, col, flu = fields(10, 11 + kind, (10, 12, 8), with_collider, with_fluid)
    src = np.random.default_rng(kind).standard_normal((8, 12, 10))
    f = (src, u, v, w)[kind]
    coef, dt = 0.07, 1.0 / 90.0
    markers = np.zeros(f.shape, dtype=np.int8)
    A = np.zeros(f.shape, dtype=ROW_DTYPE)
    b = np.zeros(f.shape)
    colc = None if col is None else np.ascontiguousarray(col)
    fluc = None if flu is None else np.ascontiguousarray(flu)
    assert hc.hc_diffusion_system(C.byref(grid), C.c_int(kind), dp(f), C.c_double(coef), C.c_double(dt), C.c_int(boundary_type),
                                  dp(colc) if colc is not None else None, dp(fluc) if fluc is not None else None,
                                  markers.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.POINTER(Row)), dp(b)) == 0
    # the oracle's solve of the same array = ICCG(100, kEpsilonD) on exactly this system
    x, _ = port.iccg_solve(A, b, 100, 2.220446049250313e-16)
    if kind == 0:
        want, _, want_markers = port.diffusion_solve_cells(grid, f, coef, dt, boundary_type, col, flu, want_markers=True)
        assert np.array_equal(markers, want_markers)
    else:
        want = port.diffusion_solve_faces(grid, u, v, w, coef, dt, boundary_type, col, flu)[kind - 1]
    assert np.array_equal(x, want), "max diff %g" % np.abs(x - want).max()
