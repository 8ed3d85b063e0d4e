"""GPU parity, seam 3: hash-grid binning (bit-exact) and the device-resident PIC / FLIP / APIC step.

Tolerances (BASELINE.json north_star): binning / cell indexing bit-exact; pressure within 1e-6 relative L-inf;
particle positions after 10 steps within 1e-5 of the domain size.  Stage outputs without reductions (markers,
particle SDF, emitted lattice) are additionally required to be bit-exact; P2G differs from the reference's serial
particle-order sum only by rounding (<= 1e-12 relative)."""
import numpy as np
import pytest

import golden_cases as gc
import scenes
from gpu_common import checker, jetcu_or_fail
from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jc():
    return jetcu_or_fail()


@pytest.fixture(scope="module")
def oracle():
    return checker()


def rel(a, b):
    return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-300)


def test_golden_bin_points(jc):
    gc.bin_points(jc)


@pytest.mark.parametrize("n,res,spacing", [(5000, (8, 8, 8), 0.11), (20000, (64, 64, 64), 0.013), (3000, (4, 2, 3), 0.5),
                                           (1, (4, 4, 4), 0.3), (300000, (64, 64, 64), 0.0331)])
def test_bin_points_bit_exact(jc, oracle, n, res, spacing):
    rng = np.random.default_rng(n)
    pts = rng.uniform(-0.3, 1.3, (n, 3))
    pts[: min(10, n)] = 0.0
    if n > 20:
        pts[10:20] = np.array([spacing, 2 * spacing, 3 * spacing])  # exactly on bucket boundaries
    r = jc.bin_points(pts, res, spacing)
    o = oracle.bin_points(pts, res, spacing)
    for k in ("keys_unsorted", "keys_sorted", "start_table", "end_table"):
        assert np.array_equal(r[k], o[k]), k
    # stable order == PointHashGridSearcher3 (serial) bucket order: ascending original index inside a bucket
    bucket, rank = oracle.bin_points_serial(pts, res, spacing)
    si, ks, st = r["sorted_indices"], r["keys_sorted"], r["start_table"]
    assert np.array_equal(bucket, r["keys_unsorted"])
    assert np.array_equal(rank[si], np.arange(len(si), dtype=np.uint64) - st[ks])


def test_bin_points_empty(jc, oracle):
    pts = np.zeros((0, 3))
    r = jc.bin_points(pts, (4, 4, 4), 0.1)
    o = oracle.bin_points(pts, (4, 4, 4), 0.1)
    assert np.array_equal(r["start_table"], o["start_table"]) and np.array_equal(r["end_table"], o["end_table"])


@pytest.mark.parametrize("mode", [PIC, FLIP, APIC])
def test_golden_hybrid(jc, mode):
    gc.hybrid(jc.hybrid, mode, exact_bits=False, tol_grid=1e-12)


def test_emit_box_bit_exact(jc, oracle):
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=PIC, mgpcg=MgpcgParams.make(max_levels=3))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    v0 = (0.1, -0.2, 0.3)
    assert a.emit_box((0.1, 0, 0.2), (0.5, 0.6, 1.5), 0.5 / n, v0) == b.emit_box((0.1, 0, 0.2), (0.5, 0.6, 1.5), 0.5 / n, v0)
    assert a.emit_box((0.6, 0, 0), (0.9, 0.3, 0.31), 0.7 / n) == b.emit_box((0.6, 0, 0), (0.9, 0.3, 0.31), 0.7 / n)
    for x, y in zip(a.get_particles(), b.get_particles()):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("lower,upper,spacing", [
    ((0.3, 0.1, 0.7), (0.7, 0.3, 0.9), 0.1 / 3), ((-0.5, 0.2, 0.25), (0.3, 2.0, 0.75), 0.05),
    ((0.05, 0.15, 0.35), (0.55, 0.45, 0.65), 0.025), ((0.07, 0.11, 0.13), (0.77, 0.51, 0.93), 0.01),
    ((0, 0, 0), (1, 0.9375, 1), 0.5 / 16), ((0.2, 0.2, 0.2), (0.2, 0.9, 0.9), 0.05), ((2, 2, 2), (3, 3, 3), 0.1)])
def test_emit_box_face_rounding_bit_exact(jc, oracle, lower, upper, spacing):
    """Lattice points that fall exactly on a face of the emitter box are kept or dropped by the reference depending
    on the rounding of its bounding-box / closest-point arithmetic; the device emitter replays it bit for bit."""
    prm = HybridParams.make(Grid.cube(16), mode=APIC, mgpcg=MgpcgParams.make(max_levels=3))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    assert a.emit_box(lower, upper, spacing) == b.emit_box(lower, upper, spacing)
    for x, y in zip(a.get_particles(), b.get_particles()):
        assert np.array_equal(x, y)


# n = 16, 32: h is a power of two (samplers take the exact-reciprocal shortcut, sampling.cuh jc_norm);
# n = 24, 20: h is not, so the literal FP64 / FP32 divisions of the reference run.
@pytest.mark.parametrize("mode,n,levels", [(PIC, 16, 3), (FLIP, 16, 3), (APIC, 16, 3), (FLIP, 32, 4), (APIC, 32, 4),
                                           (FLIP, 24, 3), (APIC, 24, 3), (PIC, 20, 2)])
def test_hybrid_stages_vs_oracle(jc, oracle, mode, n, levels):
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=levels),
                            pic_blending_factor=(0.1 if mode == FLIP and n == 16 else 0.0))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    pts = scenes.dam_particles(n)
    # shuffle so that original order != lattice order (exercises the original-order round trip)
    perm = np.random.default_rng(1).permutation(len(pts))
    pts = np.ascontiguousarray(pts[perm])
    vel = np.random.default_rng(2).uniform(-0.2, 0.2, pts.shape)
    a.set_particles(pts, vel); b.set_particles(pts, vel)
    dt = 1 / 60.0
    for step in range(3):
        a.begin_advance_time_step(dt); b.begin_advance_time_step(dt)
        for x, y in zip(a.get_markers(), b.get_markers()):
            assert np.array_equal(x, y)
        pa, _ = a.get_particles(); pb, _ = b.get_particles()
        if np.array_equal(pa, pb):  # identical particles (always true at step 0) => exact min => identical bits
            assert np.array_equal(a.get_sdf(), b.get_sdf()), "particle SDF must be bit-exact"
        else:  # positions differ by ulps after a step (P2G summation order), and so does the distance
            assert np.abs(a.get_sdf() - b.get_sdf()).max() <= 1e-12
        for x, y in zip(a.get_velocity(), b.get_velocity()):
            assert rel(x, y) <= 1e-12
        assert a.num_sub_steps(dt) == b.num_sub_steps(dt)
        assert abs(a.cfl(dt) - b.cfl(dt)) <= 1e-12 * max(abs(b.cfl(dt)), 1e-30)
        a.compute_external_forces(dt); b.compute_external_forces(dt)
        ra, rb = a.compute_pressure(dt), b.compute_pressure(dt)
        if rb.true_residual_l2 <= 1e-7 * rb.b_l2:
            assert rel(a.get_pressure(), b.get_pressure()) <= 1e-6
            assert abs(int(ra.iterations) - int(rb.iterations)) <= 1
        a.compute_advection(dt); b.compute_advection(dt)
        pa, va = a.get_particles(); pb, vb = b.get_particles()
        assert np.abs(pa - pb).max() <= 1e-5
        assert np.abs(va - vb).max() <= 1e-4 * max(np.abs(vb).max(), 1.0)


@pytest.mark.parametrize("mode", [APIC, FLIP])
def test_append_particles_keeps_device_state(jc, oracle, mode):
    """A continuous emitter between steps (ParticleSystemData3::addParticles): jetcu_hybrid_append_particles uploads the
    tail only; the old particles keep their state — for APIC the affine vectors, which `_cX.resize(n)` preserves
    (src/jet/apic_solver3.cpp:36-38) — and the new ones start from zero affine vectors."""
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=3))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    assert a.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n) == b.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n)
    dt = 1 / 60.0
    rng = np.random.default_rng(9)
    for step in range(4):
        a.step(dt); b.step(dt)
        new_p = np.ascontiguousarray(rng.uniform((0.6, 0.5, 0.2), (0.8, 0.7, 0.8), (300, 3)))
        new_v = np.ascontiguousarray(np.tile((0.0, -1.0, 0.0), (300, 1)))
        a.append_particles(new_p, new_v); b.append_particles(new_p, new_v)
        assert a.num_particles() == b.num_particles()
        if mode == APIC:
            for x, y in zip(a.get_affine(), b.get_affine()):
                assert np.abs(x - y).max() <= 1e-6 * max(np.abs(y).max(), 1.0)
                assert not x[-300:].any() and np.abs(x[:-300]).max() > 0.0
    a.step(dt); b.step(dt)
    (pa, va), (pb, vb) = a.get_particles(), b.get_particles()
    assert np.abs(pa - pb).max() <= 1e-5 and np.abs(va - vb).max() <= 1e-4 * max(np.abs(vb).max(), 1.0)
    a.close(); b.close()


@pytest.mark.parametrize("mode", [FLIP, APIC])
def test_steps_with_default_iccg_compressed(jc, oracle, mode):
    """hybrid_liquid_sim's own solver configuration (src/examples/hybrid_liquid_sim/main.cpp:219): the default FdmIccgSolver3 and
    setUseCompressedLinearSystem(true).  Five sub-steps of a dam break: positions within 1e-5 of the domain size."""
    n = 24
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.iccg(100, 1e-6, compressed=True))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    assert a.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n) == b.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n)
    for step in range(5):
        sa, sb = a.step(1 / 60.0), b.step(1 / 60.0)
        assert abs(int(sa.pcg_iterations) - int(sb.pcg_iterations)) <= 1
        pa, _ = a.get_particles(); pb, _ = b.get_particles()
        assert np.abs(pa - pb).max() <= 1e-5, "step %d: %g" % (step, np.abs(pa - pb).max())
    a.close(); b.close()


def test_particle_order_round_trip(jc):
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_levels=3))
    h = jc.hybrid(prm)
    rng = np.random.default_rng(3)
    pts = rng.uniform(0.05, 0.95, (5000, 3)); vel = rng.uniform(-1, 1, (5000, 3))
    h.set_particles(pts, vel)
    h.transfer_particles_to_grids()  # sorts the particles on the device
    p2, v2 = h.get_particles()
    assert np.array_equal(p2, pts) and np.array_equal(v2, vel)


def test_step_is_deterministic(jc):
    n = 32
    outs = []
    for _ in range(2):
        prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=4))
        h = jc.hybrid(prm)
        h.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n)
        for _ in range(3):
            h.step(1 / 60.0)
        outs.append(h.get_particles())
        h.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("mode", [FLIP, APIC])
def test_ten_steps_config1_64(jc, oracle, mode):
    """BASELINE config 1 (FLIP dam break at 64^3, 516 516 particles, MGPCG 6 levels) — 10 sub-steps:
    positions within 1e-5 of the domain size; per-solve iteration counts follow the oracle."""
    n = 64
    mg = MgpcgParams.make(max_iterations=100, max_levels=6)
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=mg)
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    na = a.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n)
    nb = b.emit_box((0, 0, 0), (0.4, 0.6, 1.0), 0.5 / n)
    assert na == nb == 516516
    dt = 1 / 60.0
    worst = 0.0
    for step in range(10):
        sa, sb = a.step(dt), b.step(dt)
        pa, _ = a.get_particles(); pb, _ = b.get_particles()
        worst = max(worst, float(np.abs(pa - pb).max()))
        assert worst <= 1e-5, "step %d: max position deviation %g (iters gpu %d / cpu %d)" % (
            step, worst, sa.pcg_iterations, sb.pcg_iterations)
