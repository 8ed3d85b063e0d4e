"""Colliders in the hybrid solver (SURVEY.md §8a-a9 with a collider, §8f-1): the C restatement
(oracle/jet_port_collider.inc) against the UNMODIFIED reference driving RigidBodyCollider3 over Box3 / Sphere3 —
collider SDF rasterisation, constrainVelocity's collider branch, extrapolateIntoCollider, the pressure system with a
moving boundary, and Collider3::resolveCollision.  Everything here is per-cell / per-particle arithmetic without
reductions except P2G (serial in both), so the port must be BIT-IDENTICAL.  CPU only; the CUDA path does not accept
colliders yet — this is the oracle it will be checked against."""
import numpy as np
import pytest

from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams, have_port, have_ref, load_port, load_ref

pytestmark = pytest.mark.skipif(not (have_port() and have_ref()), reason="needs both oracle libraries")


@pytest.fixture(scope="module")
def libs():
    ref, port = load_ref(), load_port()
    ref.set_num_threads(0)
    port.set_num_threads(0)
    return ref, port


def same(a, b, what):
    assert a.shape == b.shape, what
    assert np.array_equal(a, b), "%s: max abs diff %g" % (what, float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()))


COLLIDERS = {
    "box": dict(kind="box", a=(0.55, 0.0, 0.3), b=(0.8, 0.35, 0.7)),
    "box_moving": dict(kind="box", a=(0.55, 0.0, 0.3), b=(0.8, 0.35, 0.7), linear_velocity=(-0.4, 0.0, 0.1),
                       angular_velocity=(0.0, 0.3, 0.0), friction=0.25),
    "sphere": dict(kind="sphere", a=(0.62, 0.2, 0.5), b=(0.17, 0.0, 0.0), friction=0.1),
    "box_offgrid": dict(kind="box", a=(0.513, -0.2, 0.271), b=(0.777, 0.3333, 0.68)),
}


@pytest.mark.parametrize("name", sorted(COLLIDERS))
@pytest.mark.parametrize("n", [8, 12])
def test_collider_sdf_bit_exact(libs, name, n):
    ref, port = libs
    prm = HybridParams.make(Grid.cube(n), mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    a.set_collider_shape(**COLLIDERS[name]); b.set_collider_shape(**COLLIDERS[name])
    same(a.get_collider_sdf(), b.get_collider_sdf(), "collider sdf")
    a.close(); b.close()


@pytest.mark.parametrize("name", sorted(COLLIDERS))
def test_constrain_velocity_with_collider_bit_exact(libs, name):
    ref, port = libs
    n = 12
    grid = Grid.cube(n)
    prm = HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    a.set_collider_shape(**COLLIDERS[name]); b.set_collider_shape(**COLLIDERS[name])
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, (n, n, n + 1)); v = rng.uniform(-1, 1, (n, n + 1, n)); w = rng.uniform(-1, 1, (n + 1, n, n))
    a.set_velocity(u, v, w); b.set_velocity(u, v, w)
    a.apply_boundary_condition(); b.apply_boundary_condition()
    for x, y, what in zip(a.get_velocity(), b.get_velocity(), "uvw"):
        same(x, y, "constrained " + what)
    a.close(); b.close()


@pytest.mark.parametrize("mode", [PIC, FLIP, APIC])
@pytest.mark.parametrize("name", ["box", "box_moving", "sphere"])
def test_steps_with_collider_bit_exact(libs, mode, name):
    """Dam column collapsing onto the collider: three sub-steps, every intermediate field bit-identical."""
    ref, port = libs
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=40, max_levels=3))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    a.set_collider_shape(**COLLIDERS[name]); b.set_collider_shape(**COLLIDERS[name])
    box = ((0.0, 0.0, 0.0), (0.7, 0.7, 1.0))   # overlaps the collider: resolveCollision has work from the first step on
    assert a.emit_box(box[0], box[1], 0.5 / n, (0.8, 0.0, 0.0)) == b.emit_box(box[0], box[1], 0.5 / n, (0.8, 0.0, 0.0))
    c = COLLIDERS[name]

    def inside(pos):
        if c["kind"] == "box":
            lo, hi = np.array(c["a"]), np.array(c["b"])
            return np.all((pos > lo + 1e-9) & (pos < hi - 1e-9), axis=1)
        return np.linalg.norm(pos - np.array(c["a"]), axis=1) < c["b"][0] - 1e-9

    assert inside(b.get_particles()[0]).sum() > 50
    for s in range(3):
        sa, sb = a.step(1.0 / 60.0), b.step(1.0 / 60.0)
        assert sa.pcg_iterations == sb.pcg_iterations, (s, sa.pcg_iterations, sb.pcg_iterations)
        pa, va = a.get_particles(); pb, vb = b.get_particles()
        same(pa, pb, "positions step %d" % s); same(va, vb, "velocities step %d" % s)
        for x, y, what in zip(a.get_velocity(), b.get_velocity(), "uvw"):
            same(x, y, "grid %s step %d" % (what, s))
        same(a.get_sdf(), b.get_sdf(), "particle sdf step %d" % s)
        same(a.get_pressure(), b.get_pressure(), "pressure step %d" % s)
        assert inside(pb).sum() == 0, "particles left inside the collider after resolveCollision"
    assert (b.get_collider_sdf() < 0).any()
    a.close(); b.close()


def test_stage_order_with_collider(libs):
    """Stage-wise drive (the virtuals an adapter overrides) with a collider equals the reference's own stages."""
    ref, port = libs
    n = 12
    prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=30, max_levels=2))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    a.set_collider_shape(**COLLIDERS["box_moving"]); b.set_collider_shape(**COLLIDERS["box_moving"])
    for h in (a, b):
        h.emit_box((0.0, 0.0, 0.0), (0.6, 0.6, 1.0), 0.5 / n, (0.5, 0.0, 0.0))
    dt = 1.0 / 60.0
    for h in (a, b):
        h.begin_advance_time_step(dt)
    for x, y, what in zip(a.get_velocity(), b.get_velocity(), "uvw"):
        same(x, y, "after begin " + what)
    same(a.get_sdf(), b.get_sdf(), "sdf after begin (extrapolated into the collider)")
    for h in (a, b):
        h.compute_external_forces(dt)
        h.compute_pressure(dt)
    for x, y, what in zip(a.get_velocity(), b.get_velocity(), "uvw"):
        same(x, y, "after pressure " + what)
    for h in (a, b):
        h.compute_advection(dt)
    pa, va = a.get_particles(); pb, vb = b.get_particles()
    same(pa, pb, "positions"); same(va, vb, "velocities")
    a.close(); b.close()
