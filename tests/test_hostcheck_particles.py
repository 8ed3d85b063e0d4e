"""The G2P and advection kernels' per-particle bodies (csrc/particle_ops.cuh: sample_velocity<POW2, I>, g2p_flip_particle,
apic_affine, move_particle) executed on the CPU through tests/hostcheck, against the C oracle — bit for bit (these
stages have no reductions).  Covers all four sampler instantiations (exact-reciprocal / division normalisation x 32- /
64-bit sample indices) on power-of-two and non-power-of-two spacings."""
import ctypes as C

import numpy as np
import pytest

from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams, have_port, load_port
from test_hostcheck_collider import _load_hostcheck, dp, exact

pytestmark = pytest.mark.skipif(not have_port(), reason="oracle/libjetport.so not built")
PF = C.POINTER(C.c_float)


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


def scene(port, mode, n, domain):
    grid = Grid.cube(n, domain)
    prm = HybridParams.make(grid, mode=mode, mgpcg=MgpcgParams.make(max_iterations=30, max_levels=2))
    h = port.hybrid(prm)
    h.emit_box((0.0, 0.0, 0.0), (0.55 * domain, 0.7 * domain, domain), 0.5 * domain / n, (0.7, -0.2, 0.3))
    return grid, prm, h


def variants(domain):
    return [0, 2] + ([1, 3] if domain == 1.0 else [])   # POW2 instantiations only where h is a power of two


@pytest.mark.parametrize("n,domain", [(16, 1.0), (12, 0.9)])
def test_move_particles(hc, port, n, domain):
    grid, prm, h = scene(port, PIC, n, domain)
    dt = 1.0 / 60.0
    h.step(dt)                      # a developed velocity field
    u, v, w = h.get_velocity()
    pos, vel = h.get_particles()
    pos[:50] = np.array([0.0, 0.0, 0.0]) + 1e-9   # some particles that will hit the walls
    vel[:50] = -5.0
    h.set_particles(pos, vel)
    h.move_particles(dt)
    want_p, want_v = h.get_particles()
    assert np.abs(want_p - pos).max() > 1e-4
    for var in variants(domain):
        gp, gv = pos.copy(), vel.copy()
        assert hc.hc_move_particles(C.byref(grid), dp(u), dp(v), dp(w), C.c_double(dt), C.c_double(prm.max_cfl),
                                    C.c_int(prm.closed_domain_boundary_flag), C.c_int(var), C.c_uint64(len(gp)), dp(gp), dp(gv)) == 0
        exact(gp, want_p, "moved positions (variant %d)" % var); exact(gv, want_v, "velocities after wall clamp (variant %d)" % var)
    # the fused PIC form (k_pic_move): transferFromGridsToParticles + moveParticles in one pass over the particles
    h.set_particles(pos, vel)
    h.transfer_grids_to_particles()
    h.move_particles(dt)
    want_p, want_v = h.get_particles()
    for var in variants(domain):
        gp, gv = pos.copy(), vel.copy()
        assert hc.hc_move_particles(C.byref(grid), dp(u), dp(v), dp(w), C.c_double(dt), C.c_double(prm.max_cfl),
                                    C.c_int(prm.closed_domain_boundary_flag), C.c_int(var | 4), C.c_uint64(len(gp)), dp(gp), dp(gv)) == 0
        exact(gp, want_p, "fused G2P + move positions (variant %d)" % var); exact(gv, want_v, "fused G2P + move velocities (variant %d)" % var)
    h.close()


@pytest.mark.parametrize("n,domain", [(16, 1.0), (12, 0.9)])
def test_g2p_pic(hc, port, n, domain):
    grid, prm, h = scene(port, PIC, n, domain)
    h.step(1.0 / 60.0)
    u, v, w = h.get_velocity()
    pos, _ = h.get_particles()
    h.transfer_grids_to_particles()
    _, want = h.get_particles()
    for var in variants(domain):
        got = np.zeros_like(want)
        assert hc.hc_g2p_pic(C.byref(grid), dp(u), dp(v), dp(w), C.c_int(var), C.c_uint64(len(pos)), dp(pos), dp(got)) == 0
        exact(got, want, "PIC G2P (variant %d)" % var)
    h.close()


@pytest.mark.parametrize("n,domain", [(16, 1.0), (12, 0.9)])
@pytest.mark.parametrize("blend", [0.0, 0.3])
def test_g2p_flip(hc, port, n, domain, blend):
    grid = Grid.cube(n, domain)
    prm = HybridParams.make(grid, mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=30, max_levels=2), pic_blending_factor=blend)
    h = port.hybrid(prm)
    h.emit_box((0.0, 0.0, 0.0), (0.55 * domain, 0.7 * domain, domain), 0.5 * domain / n, (0.7, -0.2, 0.3))
    dt = 1.0 / 60.0
    h.apply_boundary_condition()
    h.transfer_particles_to_grids()
    snap = [a.astype(np.float32) for a in h.get_velocity()]          # FlipSolver3::_uDelta snapshot (flip_solver3.cpp:39-47)
    h.build_signed_distance_field(); h.extrapolate_velocity_to_air(); h.apply_boundary_condition()
    h.compute_external_forces(dt); h.compute_pressure(dt)
    h.extrapolate_velocity_to_air(); h.apply_boundary_condition()
    u, v, w = h.get_velocity()
    delta = [np.ascontiguousarray(a.astype(np.float32) - s0) for a, s0 in zip((u, v, w), snap)]   # :57-70, float arithmetic
    pos, vel = h.get_particles()
    h.transfer_grids_to_particles()
    _, want = h.get_particles()
    got = vel.copy()
    fp = lambda a: a.ctypes.data_as(PF)
    assert hc.hc_g2p_flip(C.byref(grid), dp(u), dp(v), dp(w), fp(delta[0]), fp(delta[1]), fp(delta[2]), C.c_double(blend),
                          C.c_uint64(len(pos)), dp(pos), dp(got)) == 0
    exact(got, want, "FLIP G2P")
    h.close()


@pytest.mark.parametrize("n,domain", [(16, 1.0), (12, 0.9)])
def test_g2p_apic(hc, port, n, domain):
    grid, prm, h = scene(port, APIC, n, domain)
    h.step(1.0 / 60.0)
    u, v, w = h.get_velocity()
    pos, _ = h.get_particles()
    h.transfer_grids_to_particles()
    _, want_v = h.get_particles()
    want_c = h.get_affine()
    gv = np.zeros_like(want_v)
    gc = [np.zeros_like(want_v) for _ in range(3)]
    assert hc.hc_g2p_apic(C.byref(grid), dp(u), dp(v), dp(w), C.c_uint64(len(pos)), dp(pos), dp(gv), dp(gc[0]), dp(gc[1]), dp(gc[2])) == 0
    exact(gv, want_v, "APIC G2P velocity")
    for a, b, nm in zip(gc, want_c, ("cX", "cY", "cZ")):
        exact(a, b, "APIC affine " + nm)
    h.close()
