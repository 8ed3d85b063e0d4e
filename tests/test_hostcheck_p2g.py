"""The particle-to-grid transfer's arithmetic (csrc/p2g_ops.cuh: per-cell register accumulation, per-face gather of a tile's
partial sums, colour-ordered tile passes, normalisation) executed on the CPU through tests/hostcheck:
  * against the C oracle (the reference's serial particle-order sum): markers bit-exact, velocities within the
    documented 1e-12 relative (only the summation order differs);
  * z-slab decompositions (2 and 4 "ranks", each with its slab + 2 ghost planes of particles) against the undivided
    domain: BIT-IDENTICAL — the summation order is fixed by the cell and tile lattice, not by who holds the particles."""
import ctypes as C

import numpy as np
import pytest

from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams, have_port, load_port
from test_hostcheck_collider import _load_hostcheck, dp, exact

pytestmark = pytest.mark.skipif(not have_port(), reason="oracle/libjetport.so not built")
PC = C.POINTER(C.c_char)


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


def run_hc(hc, grid, apic, tz, slabs, pos, vel, aff):
    nx, ny, nz = grid.res
    u = np.zeros((nz, ny, nx + 1)); v = np.zeros((nz, ny + 1, nx)); w = np.zeros((nz + 1, ny, nx))
    um = np.zeros(u.shape, dtype=np.int8); vm = np.zeros(v.shape, dtype=np.int8); wm = np.zeros(w.shape, dtype=np.int8)
    err = C.c_int(0)
    cp = lambda a: a.ctypes.data_as(PC)
    z = np.zeros_like(pos)
    cx, cy, cz = aff if aff is not None else (z, z, z)
    rc = hc.hc_p2g(C.byref(grid), C.c_int(1 if apic else 0), C.c_int(tz), C.c_int(slabs), C.c_uint64(len(pos)), dp(pos), dp(vel),
                   dp(cx), dp(cy), dp(cz), dp(u), dp(v), dp(w), cp(um), cp(vm), cp(wm), C.byref(err))
    assert rc == 0 and err.value == 0
    return (u, v, w), (um, vm, wm)


@pytest.mark.parametrize("tz", [0, 4], ids=["staggered", "window"])   # 0: k_p2g_stag (the default), 4: k_p2g_tile (A/B kernel)
@pytest.mark.parametrize("mode", [PIC, APIC])
@pytest.mark.parametrize("n,domain,jitter", [(16, 1.0, 0.0), (12, 0.9, 0.37), (20, 1.0, 0.49)])
def test_p2g_against_oracle_and_between_decompositions(hc, port, mode, n, domain, jitter, tz):
    grid = Grid.cube(n, domain)
    prm = HybridParams.make(grid, mode=mode, mgpcg=MgpcgParams.make(max_iterations=20, max_levels=2))
    h = port.hybrid(prm)
    h.emit_box((0.0, 0.0, 0.0), (0.6 * domain, 0.75 * domain, domain), 0.5 * domain / n, (0.7, -0.2, 0.3))
    pos, vel = h.get_particles()
    rng = np.random.default_rng(n)
    if jitter:   # off-lattice positions (generic weights) and a rough velocity field with a wide dynamic range
        pos = np.ascontiguousarray(np.clip(pos + rng.uniform(-jitter, jitter, pos.shape) * domain / n, 0.0, domain))
        vel = vel * rng.uniform(0.0, 2.0, vel.shape) * np.where(rng.uniform(size=vel.shape) < 0.05, 1e-9, 1.0)
    aff = None
    if mode == APIC:
        h.set_particles(pos, vel)
        h.step(1.0 / 120.0)           # gives the particles non-trivial affine vectors
        pos, vel = h.get_particles()
        aff = h.get_affine()
    else:
        h.set_particles(pos, vel)
    h.transfer_particles_to_grids()
    want = h.get_velocity()
    want_m = h.get_markers()
    got, got_m = run_hc(hc, grid, mode == APIC, tz, 1, pos, vel, aff)
    scale = max(float(np.abs(a).max()) for a in want)
    for a, b, am, bm, nm in zip(got, want, got_m, want_m, "uvw"):
        exact(am, bm, "marker " + nm)
        assert np.abs(a - b).max() <= 1e-12 * scale, "%s: %g" % (nm, np.abs(a - b).max() / scale)
    # z-slab decompositions: every rank sums the ghost cells along, tile by tile in the same order -> the same bits
    for slabs in (2, 4):
        if n % slabs == 0 and n // slabs >= 4:
            got_s, got_m_s = run_hc(hc, grid, mode == APIC, tz, slabs, pos, vel, aff)
            for a, b, am, bm, nm in zip(got_s, got, got_m_s, got_m, "uvw"):
                exact(a, b, "%d slabs vs one domain: %s" % (slabs, nm)); exact(am, bm, "%d slabs: marker %s" % (slabs, nm))


@pytest.mark.parametrize("mode", [PIC, APIC])
@pytest.mark.parametrize("res", [(2, 2, 2), (1, 5, 3), (3, 1, 4), (4, 3, 1), (1, 1, 1), (9, 2, 17), (17, 9, 5)])
def test_p2g_staggered_degenerate_grids(hc, port, mode, res):
    """Grids with one or two cells along an axis (a single staggered cell whose halves overlap, clamped bases on both walls
    at once) and tile-straddling sizes; particles on the domain faces and corners included."""
    rng = np.random.default_rng(sum(res) * 7 + mode)
    grid = Grid(res[0], res[1], res[2], 0.25, 0.25, 0.25, 0.5, -1.0, 0.0)
    prm = HybridParams.make(grid, mode=mode, mgpcg=MgpcgParams.make(max_iterations=5, max_levels=1))
    h = port.hybrid(prm)
    lo = np.array([grid.ox, grid.oy, grid.oz]); size = np.array(res) * 0.25
    n = 40 * res[0] * res[1] * res[2] + 30
    pos = lo + rng.uniform(0.0, 1.0, (n, 3)) * size
    pos[:24] = lo + size * rng.choice([0.0, 0.5, 1.0], (24, 3))
    vel = rng.uniform(-1.0, 1.0, (n, 3))
    h.set_particles(np.ascontiguousarray(pos), vel)
    aff = None
    if mode == APIC:
        h.step(1.0 / 240.0)           # gives the particles non-trivial affine vectors
        aff = h.get_affine()
    pos, vel = h.get_particles()
    assert np.isfinite(pos).all() and np.isfinite(vel).all()
    h.transfer_particles_to_grids()
    want, want_m = h.get_velocity(), h.get_markers()
    got, got_m = run_hc(hc, grid, mode == APIC, 0, 1, pos, vel, aff)
    scale = max(float(np.abs(a).max()) for a in want)
    for a, b, am, bm, nm in zip(got, want, got_m, want_m, "uvw"):
        exact(am, bm, "marker " + nm)
        assert np.abs(a - b).max() <= 1e-12 * scale, "%s: %g" % (nm, np.abs(a - b).max() / scale)
    h.close()
