"""Colliders in the hybrid solver (SURVEY.md §8a-a9 with a collider, §8f-1): the C restatement
(oracle/jet_port_collider.inc) against the UNMODIFIED reference driving RigidBodyCollider3 over Box3 / Sphere3 /
Cylinder3 / Plane3 (plain or normal-flipped) or an ImplicitSurfaceSet3 of them (Bvh3 traversal included) —
collider SDF rasterisation, constrainVelocity's collider branch, extrapolateIntoCollider, the pressure system with a
moving boundary, and Collider3::resolveCollision.  Everything here is per-cell / per-particle arithmetic without
reductions except P2G (serial in both), so the port must be BIT-IDENTICAL.  CPU only; this pins the oracle the CUDA
path is checked against (tests/test_hostcheck_collider.py on the CPU, tests/test_gpu_collider.py on the GPU)."""
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
    "cylinder": dict(kind="cylinder", a=(0.62, 0.25, 0.5), b=(0.15, 0.5, 0.0), friction=0.2),
    "plane": dict(kind="plane", a=(0.0, 0.13, 0.0), b=(0.0, 1.0, 0.0)),                      # a raised floor
    "plane_tilted": dict(kind="plane", a=(0.5, 0.2, 0.5), b=(0.6, 0.8, 0.0), friction=0.3),
    "tank": dict(kind="sphere", a=(0.5, 0.5, 0.5), b=(0.45, 0.0, 0.0), flipped=True),         # hybrid_liquid_sim example 6
}

# ImplicitSurfaceSet3 colliders: (shapes, kwargs).  "cylinders" follows hybrid_liquid_sim example 3 (three pillars in the path of a dam break).
SETS = {
    "cylinders": ([dict(kind="cylinder", a=(1.0 / 3, 0.25, 0.2), b=(0.08, 0.5)),
                   dict(kind="cylinder", a=(0.5, 0.25, 0.5), b=(0.08, 0.5)),
                   dict(kind="cylinder", a=(2.0 / 3, 0.25, 0.8), b=(0.08, 0.5))], dict()),
    "mixed": ([dict(kind="box", a=(0.55, 0.0, 0.3), b=(0.8, 0.35, 0.7)),
               dict(kind="sphere", a=(0.3, 0.3, 0.6), b=(0.12, 0.0, 0.0)),
               dict(kind="cylinder", a=(0.3, 0.2, 0.2), b=(0.1, 0.4)),
               dict(kind="plane", a=(0.0, 0.07, 0.0), b=(0.0, 1.0, 0.0))],
              dict(linear_velocity=(0.1, 0.0, -0.2), angular_velocity=(0.0, 0.2, 0.1), friction=0.15)),
    "single_in_set": ([dict(kind="sphere", a=(0.62, 0.2, 0.5), b=(0.17, 0.0, 0.0))], dict(friction=0.1)),
    "symmetric_spheres": ([dict(kind="sphere", a=(0.25, 0.25, 0.5), b=(0.125, 0.0, 0.0)),         # many exact ties: the
                           dict(kind="sphere", a=(0.75, 0.25, 0.5), b=(0.125, 0.0, 0.0)),         # Bvh3 traversal order
                           dict(kind="sphere", a=(0.5, 0.25, 0.25), b=(0.125, 0.0, 0.0)),         # decides which member
                           dict(kind="sphere", a=(0.5, 0.25, 0.75), b=(0.125, 0.0, 0.0)),         # answers
                           dict(kind="box", a=(0.4375, 0.0, 0.4375), b=(0.5625, 0.5, 0.5625))], dict()),
    "planes_only": ([dict(kind="plane", a=(0.0, 0.1, 0.0), b=(0.0, 1.0, 0.0)),
                     dict(kind="plane", a=(0.9, 0.0, 0.0), b=(-1.0, 0.0, 0.0))], dict()),          # empty BVH
}


def install(h, name):
    if name in COLLIDERS:
        h.set_collider_shape(**COLLIDERS[name])
    else:
        shapes, kw = SETS[name]
        h.set_collider_shapes(shapes, True, **kw)


ALL = sorted(COLLIDERS) + sorted(SETS)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("n", [8, 12])
def test_collider_sdf_bit_exact(libs, name, n):
    ref, port = libs
    prm = HybridParams.make(Grid.cube(n), mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    install(a, name); install(b, name)
    same(a.get_collider_sdf(), b.get_collider_sdf(), "collider sdf")
    a.close(); b.close()


@pytest.mark.parametrize("name", ALL)
def test_constrain_velocity_with_collider_bit_exact(libs, name):
    ref, port = libs
    n = 12
    grid = Grid.cube(n)
    prm = HybridParams.make(grid, mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    install(a, name); install(b, name)
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, (n, n, n + 1)); v = rng.uniform(-1, 1, (n, n + 1, n)); w = rng.uniform(-1, 1, (n + 1, n, n))
    a.set_velocity(u, v, w); b.set_velocity(u, v, w)
    a.apply_boundary_condition(); b.apply_boundary_condition()
    for x, y, what in zip(a.get_velocity(), b.get_velocity(), "uvw"):
        same(x, y, "constrained " + what)
    a.close(); b.close()


STEP_CASES = ["box", "box_moving", "sphere", "cylinder", "plane_tilted", "tank", "cylinders", "mixed", "symmetric_spheres"]


@pytest.mark.parametrize("mode", [PIC, FLIP, APIC])
@pytest.mark.parametrize("name", STEP_CASES)
def test_steps_with_collider_bit_exact(libs, mode, name):
    """Dam column collapsing onto the collider: three sub-steps, every intermediate field bit-identical."""
    ref, port = libs
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=40, max_levels=3))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    install(a, name); install(b, name)
    box = ((0.0, 0.0, 0.0), (0.7, 0.7, 1.0))   # overlaps the collider: resolveCollision has work from the first step on
    assert a.emit_box(box[0], box[1], 0.5 / n, (0.8, 0.0, 0.0)) == b.emit_box(box[0], box[1], 0.5 / n, (0.8, 0.0, 0.0))
    csdf = b.get_collider_sdf()

    def inside(pos):   # strictly inside the collider, judged on the (trilinearly sampled) rasterised SDF
        from scipy.ndimage import map_coordinates
        idx = (pos * n - 0.5).T[::-1]
        interior = np.all((pos >= 0.5 / n) & (pos <= 1.0 - 0.5 / n), axis=1)   # no extrapolated samples
        return interior & (map_coordinates(csdf, idx, order=1, mode="nearest") < -0.02)

    assert inside(b.get_particles()[0]).sum() > 20
    for s in range(3):
        sa, sb = a.step(1.0 / 60.0), b.step(1.0 / 60.0)
        assert sa.pcg_iterations == sb.pcg_iterations, (s, sa.pcg_iterations, sb.pcg_iterations)
        pa, va = a.get_particles(); pb, vb = b.get_particles()
        same(pa, pb, "positions step %d" % s); same(va, vb, "velocities step %d" % s)
        for x, y, what in zip(a.get_velocity(), b.get_velocity(), "uvw"):
            same(x, y, "grid %s step %d" % (what, s))
        same(a.get_sdf(), b.get_sdf(), "particle sdf step %d" % s)
        same(a.get_pressure(), b.get_pressure(), "pressure step %d" % s)
        if name != "mixed":   # overlapping members: the set projects onto the NEAREST member's surface, which may lie inside another
            assert inside(pb).sum() == 0, "particles left inside the collider after resolveCollision"
    assert (csdf < 0).any()
    a.close(); b.close()


@pytest.mark.parametrize("name", ["box_moving", "mixed"])
def test_stage_order_with_collider(libs, name):
    """Stage-wise drive (the virtuals an adapter overrides) with a collider equals the reference's own stages."""
    ref, port = libs
    n = 12
    prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=30, max_levels=2))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    install(a, name); install(b, name)
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


EMITTERS = {   # the emitters of src/examples/hybrid_liquid_sim (half-space + sphere; two boxes) and every other kind
    "sphere": [dict(kind="sphere", a=(0.5, 0.5, 0.5), b=(0.3, 0.0, 0.0))],
    "half_space": [dict(kind="plane", a=(0.0, 0.25, 0.0), b=(0.0, 1.0, 0.0))],
    "cylinder": [dict(kind="cylinder", a=(0.4, 0.3, 0.6), b=(0.2, 0.5))],
    "box": [dict(kind="box", a=(0.0, 0.0, 0.0), b=(0.5, 0.9375, 1.0))],
    "two_boxes": [dict(kind="box", a=(0.0, 0.0, 0.0), b=(0.17 + 0.001, 0.25 + 0.001, 0.375 + 0.001)),
                  dict(kind="box", a=(0.83 - 0.001, 0.0, 0.125 - 0.001), b=(1.17 + 0.001, 0.25 + 0.001, 0.75 + 0.001))],
    "mixed": [dict(kind="sphere", a=(0.5, 0.5, 0.5), b=(0.2, 0.0, 0.0)), dict(kind="plane", a=(0.0, 0.25, 0.0), b=(0.0, 1.0, 0.0)),
              dict(kind="cylinder", a=(0.8, 0.5, 0.2), b=(0.1, 0.6))],
    "outside_of_sphere": [dict(kind="sphere", a=(0.5, 0.5, 0.5), b=(0.3, 0.0, 0.0), flipped=True)],
}


@pytest.mark.parametrize("lattice", ["grid", "bcc"])      # GridPointGenerator3 / BccLatticePointGenerator (the default)
@pytest.mark.parametrize("name", sorted(EMITTERS))
@pytest.mark.parametrize("n,domain", [(16, 1.0), (10, 0.9)])
def test_emit_shapes_bit_exact(libs, name, n, domain, lattice):
    """VolumeParticleEmitter3 (one shot, jitter 0, GridPointGenerator3, maxRegion = domain) over analytic surfaces: the
    restatement emits the same points in the same order as the unmodified reference."""
    ref, port = libs
    prm = HybridParams.make(Grid.cube(n, domain), mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = ref.hybrid(prm), port.hybrid(prm)
    spacing = 0.5 * domain / n
    na, nb = (s.emit_shapes(EMITTERS[name], spacing, (0.1, 0.0, -0.2), lattice=lattice) for s in (a, b))
    assert na == nb and na > 100
    (pa, va), (pb, vb) = a.get_particles(), b.get_particles()
    same(pa, pb, "emitted positions"); same(va, vb, "emitted velocities")
    # a second emitter appends (ParticleEmitterSet3: one after the other)
    assert a.emit_shapes(EMITTERS["sphere"], spacing, lattice=lattice) == b.emit_shapes(EMITTERS["sphere"], spacing, lattice=lattice)
    same(a.get_particles()[0], b.get_particles()[0], "positions after the second emitter")
    a.close(); b.close()
