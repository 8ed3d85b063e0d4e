"""Deterministic synthetic inputs for the parity tests and bench.py.

All arrays follow Jet's Array3 layout: numpy shape (nz, ny, nx), x fastest.
"""
import numpy as np

from oracle_api import Grid, ROW_DTYPE, size3


def cell_centers(grid):
    nx, ny, nz = grid.res
    x = grid.ox + (np.arange(nx) + 0.5) * grid.hx
    y = grid.oy + (np.arange(ny) + 0.5) * grid.hy
    z = grid.oz + (np.arange(nz) + 0.5) * grid.hz
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    return X, Y, Z


def box_sdf(X, Y, Z, lower, upper):
    """Signed distance to an axis-aligned box (negative inside)."""
    dx = np.maximum(lower[0] - X, X - upper[0])
    dy = np.maximum(lower[1] - Y, Y - upper[1])
    dz = np.maximum(lower[2] - Z, Z - upper[2])
    outside = np.sqrt(np.maximum(dx, 0) ** 2 + np.maximum(dy, 0) ** 2 + np.maximum(dz, 0) ** 2)
    inside = np.minimum(np.maximum(dx, np.maximum(dy, dz)), 0)
    return outside + inside


def face_velocity(grid, value=(0.0, 0.0, 0.0), wall_zero=True):
    u = np.full(size3(grid.usize), value[0])
    v = np.full(size3(grid.vsize), value[1])
    w = np.full(size3(grid.wsize), value[2])
    if wall_zero:
        u[:, :, 0] = 0; u[:, :, -1] = 0
        v[:, 0, :] = 0; v[:, -1, :] = 0
        w[0, :, :] = 0; w[-1, :, :] = 0
    return u, v, w


BOX_COLLIDER = ((0.15, 0.0, 0.3), (0.3, 0.25, 0.7))


def projection_scene(n, kind="slab", collider=False):
    """Config-2 family: unit cube, h = 1/n.
    kind 'slab'  : fluidSdf = y - 0.5, v = +1 on interior faces  (oracle converges at every size)
    kind 'column': fluidSdf = max(x - 0.4, y - 0.6), v = -1      (dam column; oracle stalls at >= 128^3)
    """
    grid = Grid.cube(n)
    X, Y, Z = cell_centers(grid)
    if kind == "slab":
        fluid = Y - 0.5
        vel = (0.0, 1.0, 0.0)
    elif kind == "column":
        fluid = np.maximum(X - 0.4, Y - 0.6)
        vel = (0.0, -1.0, 0.0)
    else:
        raise ValueError(kind)
    u, v, w = face_velocity(grid, vel)
    csdf = box_sdf(X, Y, Z, *BOX_COLLIDER) if collider else None
    return dict(grid=grid, u=u, v=v, w=w, fluid_sdf=np.ascontiguousarray(fluid), collider_sdf=csdf)


def random_system(nx, ny, nz, seed=0, spd=True):
    """A random 7-point system.  spd=True gives a diagonally dominant M-matrix (like the Poisson system);
    spd=False mirrors the reference's benchmark fixture: U(0,1) entries
    (src/tests/time_perf_tests/fdm_linear_systems_tests.cpp:47-70)."""
    rng = np.random.default_rng(seed)
    A = np.zeros((nz, ny, nx), dtype=ROW_DTYPE)
    if spd:
        r = -rng.uniform(0.5, 1.5, size=(nz, ny, nx))
        up = -rng.uniform(0.5, 1.5, size=(nz, ny, nx))
        fr = -rng.uniform(0.5, 1.5, size=(nz, ny, nx))
        r[:, :, -1] = 0; up[:, -1, :] = 0; fr[-1, :, :] = 0
        c = -(r + up + fr)
        c[:, :, 1:] += -r[:, :, :-1]
        c[:, 1:, :] += -up[:, :-1, :]
        c[1:, :, :] += -fr[:-1, :, :]
        c += rng.uniform(0.01, 0.1, size=c.shape)
        A["center"], A["right"], A["up"], A["front"] = c, r, up, fr
    else:
        for f in ("center", "right", "up", "front"):
            A[f] = rng.uniform(0, 1, size=(nz, ny, nx))
    x = rng.uniform(-1, 1, size=(nz, ny, nx))
    b = rng.uniform(-1, 1, size=(nz, ny, nx))
    return A, x, b


def poisson_levels(n, levels, seed=None):
    """Neumann Poisson matrices on n^3, n/2^3, ... like the reference's MG test fixture
    (src/tests/unit_tests/fdm_mgpcg_solver3_tests.cpp:13-60 builds the same stencil per level)."""
    out = []
    for l in range(levels):
        m = n >> l
        A = np.zeros((m, m, m), dtype=ROW_DTYPE)
        c = np.zeros((m, m, m))
        r = np.full((m, m, m), -1.0); r[:, :, -1] = 0
        up = np.full((m, m, m), -1.0); up[:, -1, :] = 0
        fr = np.full((m, m, m), -1.0); fr[-1, :, :] = 0
        c[:, :, :-1] += 1; c[:, :, 1:] += 1
        c[:, :-1, :] += 1; c[:, 1:, :] += 1
        c[:-1, :, :] += 1; c[1:, :, :] += 1
        A["center"], A["right"], A["up"], A["front"] = c, r, up, fr
        out.append(A)
    return out


def dam_particles(n, lower=(0.0, 0.0, 0.0), upper=(0.4, 0.6, 1.0), domain=1.0):
    """Lattice the reference's VolumeParticleEmitter3 + GridPointGenerator3 produce for a box with spacing h/2,
    in k,j,i order (src/jet/grid_point_generator3.cpp:12-37)."""
    h = domain / n
    s = 0.5 * h
    def axis(lo, hi):
        pts = []
        i = 0
        while i * s <= (hi - lo):
            pts.append(i * s + lo)
            i += 1
        return np.array(pts)
    xs, ys, zs = axis(lower[0], upper[0]), axis(lower[1], upper[1]), axis(lower[2], upper[2])
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    return np.ascontiguousarray(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1))
