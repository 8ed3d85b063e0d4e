"""GridBackwardEulerDiffusionSolver3: the C restatement (oracle/jet_port_diffusion.inc) against the UNMODIFIED
reference (oracle/_ref), bit for bit — markers, the three face solves, the cell-centred solve — and GridFluidSolver3::
computeViscosity inside the hybrid step.  CPU only."""
import numpy as np
import pytest

from oracle_api import FLIP, PIC, Grid, HybridParams, MgpcgParams, have_port, have_ref, load_port, load_ref

pytestmark = pytest.mark.skipif(not (have_port() and have_ref()), reason="needs both oracle libraries")


@pytest.fixture(scope="module")
def libs():
    return load_port(), load_ref()


def fields(n, seed, shape3, with_collider, with_fluid):
    """Random face velocities, a spherical collider SDF and a tilted-plane fluid SDF on an n x (n+2) x (n-2) grid."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape3
    h = 1.0 / n
    grid = Grid.cube(n)
    grid.nx, grid.ny, grid.nz = nx, ny, nz
    z, y, x = np.meshgrid((np.arange(nz) + 0.5) * h, (np.arange(ny) + 0.5) * h, (np.arange(nx) + 0.5) * h, indexing="ij")
    col = np.sqrt((x - 0.45) ** 2 + (y - 0.3) ** 2 + (z - 0.4) ** 2) - 0.17 if with_collider else None
    flu = (y + 0.3 * x - 0.2 * z) - 0.62 if with_fluid else None
    u = rng.standard_normal((nz, ny, nx + 1)); v = rng.standard_normal((nz, ny + 1, nx)); w = rng.standard_normal((nz + 1, ny, nx))
    return grid, u, v, w, col, flu


@pytest.mark.parametrize("boundary_type", [0, 1])
@pytest.mark.parametrize("n,shape3,with_collider,with_fluid",
                         [(12, (12, 14, 10), True, True), (10, (10, 10, 10), False, True), (9, (9, 8, 11), True, False)])
def test_face_solve_bit_exact(libs, n, shape3, with_collider, with_fluid, boundary_type):
    port, ref = libs
    grid, u, v, w, col, flu = fields(n, 7 + n, shape3, with_collider, with_fluid)
    coef, dt = 0.05, 1.0 / 60.0
    a = port.diffusion_solve_faces(grid, u, v, w, coef, dt, boundary_type, col, flu)
    b = ref.diffusion_solve_faces(grid, u, v, w, coef, dt, boundary_type, col, flu)
    for q in range(3):
        assert np.array_equal(a[q], b[q]), "component %d differs: %g" % (q, np.abs(a[q] - b[q]).max())
        assert np.abs(a[q] - (u, v, w)[q]).max() > 1e-3           # something was diffused
    assert int(a[3][2].iterations) == int(b[3][2].iterations)      # the reference exposes its last (w) solve only
    assert a[3][2].last_residual == b[3][2].last_residual


@pytest.mark.parametrize("boundary_type", [0, 1])
def test_cell_solve_and_markers_bit_exact(libs, boundary_type):
    port, ref = libs
    grid, u, v, w, col, flu = fields(12, 3, (12, 14, 10), True, True)
    src = np.random.default_rng(5).standard_normal((10, 14, 12))
    a = port.diffusion_solve_cells(grid, src, 0.2, 0.01, boundary_type, col, flu, want_markers=True)
    b = ref.diffusion_solve_cells(grid, src, 0.2, 0.01, boundary_type, col, flu, want_markers=True)
    assert np.array_equal(a[2], b[2])
    assert set(np.unique(a[2])) == {0, 1, 2}                      # kFluid, kAir, kBoundary all present
    assert np.array_equal(a[0], b[0])
    assert int(a[1].iterations) == int(b[1].iterations)


@pytest.mark.parametrize("mode", [PIC, FLIP])
def test_hybrid_steps_with_viscosity_bit_exact(libs, mode):
    """GridFluidSolver3::computeViscosity (grid_fluid_solver3.cpp:231-241) inside onAdvanceTimeStep, with a box collider."""
    port, ref = libs
    n = 12
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=40, max_levels=2))
    hs = [lib.hybrid(prm) for lib in (port, ref)]
    for h in hs:
        h.emit_box((0.0, 0.0, 0.0), (0.5, 0.7, 1.0), 0.5 / n, (0.4, -0.1, 0.2))
        h.set_collider_shape("box", (0.6, 0.0, 0.3), (0.8, 0.3, 0.7))
        h.set_viscosity(0.08)
    for step in range(3):
        for h in hs:
            h.step(1.0 / 120.0)
        (pa, va), (pb, vb) = hs[0].get_particles(), hs[1].get_particles()
        assert np.array_equal(pa, pb) and np.array_equal(va, vb), "step %d" % step
        for a, b in zip(hs[0].get_velocity(), hs[1].get_velocity()):
            assert np.array_equal(a, b)
    # and it matters: the same scene without viscosity ends elsewhere
    h0 = port.hybrid(prm)
    h0.emit_box((0.0, 0.0, 0.0), (0.5, 0.7, 1.0), 0.5 / n, (0.4, -0.1, 0.2))
    h0.set_collider_shape("box", (0.6, 0.0, 0.3), (0.8, 0.3, 0.7))
    for step in range(3):
        h0.step(1.0 / 120.0)
    assert np.abs(h0.get_particles()[0] - hs[0].get_particles()[0]).max() > 1e-6
    for h in hs + [h0]:
        h.close()
