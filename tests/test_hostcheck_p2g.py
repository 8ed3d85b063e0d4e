"""The particle-to-grid transfer's arithmetic (csrc/p2g_ops.cuh: per-cell register accumulation, 128-bit fixed-point
conversion, shared-limb accumulation in BOTH representations, tile flush, normalisation) executed on the CPU through
tests/hostcheck:
  * against the C oracle (the reference's serial particle-order sum): markers bit-exact, velocities within the
    documented 1e-12 relative (only the summation order differs);
  * carry-free (26-bit redundant limbs) against the default (32-bit limbs with carry chains): BIT-IDENTICAL — both add the
    same per-cell partial sums as exact integers, only the shared-memory representation differs.  This is the check the
    experimental JETCU_P2G_KERNEL=carryfree variant needs before its first GPU run."""
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


def run_hc(hc, grid, apic, cf, pos, vel, aff):
    nx, ny, nz = grid.res
    u = np.zeros((nz, ny, nx + 1)); v = np.zeros((nz, ny + 1, nx)); w = np.zeros((nz + 1, ny, nx))
    um = np.zeros(u.shape, dtype=np.int8); vm = np.zeros(v.shape, dtype=np.int8); wm = np.zeros(w.shape, dtype=np.int8)
    err = C.c_int(0)
    cp = lambda a: a.ctypes.data_as(PC)
    z = np.zeros_like(pos)
    cx, cy, cz = aff if aff is not None else (z, z, z)
    rc = hc.hc_p2g(C.byref(grid), C.c_int(1 if apic else 0), C.c_int(1 if cf else 0), C.c_uint64(len(pos)), dp(pos), dp(vel),
                   dp(cx), dp(cy), dp(cz), dp(u), dp(v), dp(w), cp(um), cp(vm), cp(wm), C.byref(err))
    assert rc == 0 and err.value == 0
    return (u, v, w), (um, vm, wm)


@pytest.mark.parametrize("mode", [PIC, APIC])
@pytest.mark.parametrize("n,domain,jitter", [(16, 1.0, 0.0), (12, 0.9, 0.37), (20, 1.0, 0.49)])
def test_p2g_against_oracle_and_between_variants(hc, port, mode, n, domain, jitter):
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
    got, got_m = run_hc(hc, grid, mode == APIC, False, pos, vel, aff)
    got_cf, got_m_cf = run_hc(hc, grid, mode == APIC, True, pos, vel, aff)
    scale = max(float(np.abs(a).max()) for a in want)
    for a, b, am, bm, nm in zip(got, want, got_m, want_m, "uvw"):
        exact(am, bm, "marker " + nm)
        assert np.abs(a - b).max() <= 1e-12 * scale, "%s: %g" % (nm, np.abs(a - b).max() / scale)
    for a, b, am, bm, nm in zip(got_cf, got, got_m_cf, got_m, "uvw"):
        exact(a, b, "carry-free vs default " + nm)
        exact(am, bm, "carry-free marker " + nm)
    h.close()


def test_p2g_empty_and_single_particle(hc, port):
    grid = Grid.cube(8)
    (u, v, w), (um, vm, wm) = run_hc(hc, grid, False, False, np.zeros((0, 3)), np.zeros((0, 3)), None)
    assert not u.any() and not um.any()
    pos = np.array([[0.31, 0.52, 0.77]]); vel = np.array([[1.0, -2.0, 0.5]])
    (u, v, w), (um, vm, wm) = run_hc(hc, grid, False, True, pos, vel, None)
    # a single particle: every touched face carries exactly the particle's velocity component (sum / weight)
    assert um.sum() == 8 and vm.sum() == 8 and wm.sum() == 8
    assert np.allclose(u[um == 1], 1.0, rtol=1e-15) and np.allclose(v[vm == 1], -2.0, rtol=1e-15) and np.allclose(w[wm == 1], 0.5, rtol=1e-15)
