"""GPU parity for analytic rigid colliders inside the device-resident solver (SURVEY.md §8a-a9 collider branch, §8f-1).

The CUDA side (csrc/collider.cuh + the k_constrain_* / k_resolve_collision kernels) mirrors the pinned oracle
(tests/test_oracle_collider.py: port == unmodified reference, bit for bit).

Expected: collider SDF, constrained velocities, markers — bit-exact (no reductions) for boxes, spheres and planes, within
1e-14 / 1e-12 where a Cylinder3 is involved (atan2 / cos / sin: device libm vs glibc); steps — positions <= 1e-5 L,
pressure <= 1e-6 relative (P2G sums in integers, the reference in particle order)."""
import numpy as np
import pytest

import golden_cases as gc
from gpu_common import checker, jetcu_or_fail
from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams
from test_oracle_collider import ALL, COLLIDERS, EMITTERS, SETS, STEP_CASES, install

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jc():
    return jetcu_or_fail()


@pytest.fixture(scope="module")
def oracle():
    return checker()


def exact(a, b, what):
    assert np.array_equal(a, b), "%s: max abs diff %g" % (what, float(np.abs(a - b).max()))


def has_cylinder(name):
    members = [COLLIDERS[name]] if name in COLLIDERS else SETS[name][0]
    return any(m["kind"] == "cylinder" for m in members)


def same(a, b, what, name, tol):
    if has_cylinder(name):
        assert np.abs(a - b).max() <= tol, "%s: max abs diff %g" % (what, float(np.abs(a - b).max()))
    else:
        exact(a, b, what)


@pytest.mark.parametrize("name", ALL)
def test_collider_sdf_bit_exact(jc, oracle, name):
    n = 12
    prm = HybridParams.make(Grid.cube(n), mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    install(a, name); install(b, name)
    same(a.get_collider_sdf(), b.get_collider_sdf(), "collider sdf", name, 1e-14)
    a.close(); b.close()


@pytest.mark.parametrize("name", ALL)
def test_constrain_velocity_bit_exact(jc, oracle, name):
    n = 12
    prm = HybridParams.make(Grid.cube(n), mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    install(a, name); install(b, name)
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, (n, n, n + 1)); v = rng.uniform(-1, 1, (n, n + 1, n)); w = rng.uniform(-1, 1, (n + 1, n, n))
    a.set_velocity(u, v, w); b.set_velocity(u, v, w)
    a.apply_boundary_condition(); b.apply_boundary_condition()
    for x, y, what in zip(a.get_velocity(), b.get_velocity(), "uvw"):
        same(x, y, "constrained " + what, name, 1e-12)
    a.close(); b.close()


@pytest.mark.parametrize("mode", [PIC, FLIP, APIC])
@pytest.mark.parametrize("name", STEP_CASES)
def test_steps_with_collider(jc, oracle, mode, name):
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=40, max_levels=3))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    install(a, name); install(b, name)
    box = ((0.0, 0.0, 0.0), (0.7, 0.7, 1.0))
    assert a.emit_box(box[0], box[1], 0.5 / n, (0.8, 0.0, 0.0)) == b.emit_box(box[0], box[1], 0.5 / n, (0.8, 0.0, 0.0))
    for s in range(3):
        sa, sb = a.step(1.0 / 60.0), b.step(1.0 / 60.0)
        pa, va = a.get_particles(); pb, vb = b.get_particles()
        assert np.abs(pa - pb).max() <= 1e-5, "positions step %d: %g" % (s, np.abs(pa - pb).max())
        pr_a, pr_b = a.get_pressure(), b.get_pressure()
        if sb.pcg_true_residual_l2 <= 1e-8 * max(sb.pcg_b_l2, 1e-300):   # only where the oracle's own solve converged
            assert np.abs(pr_a - pr_b).max() <= 1e-6 * max(np.abs(pr_b).max(), 1e-300), "pressure step %d" % s
        if s == 0:
            same(a.get_sdf(), b.get_sdf(), "particle sdf (extrapolated into the collider) step 0", name, 1e-12)
    a.close(); b.close()


@pytest.mark.parametrize("name", sorted(gc.COLLIDER_CASES))
def test_golden_collider(jc, name):
    gc.collider(jc.hybrid, name, exact_bits=False)


@pytest.mark.parametrize("lattice", ["grid", "bcc"])
@pytest.mark.parametrize("name", sorted(EMITTERS))
def test_emit_shapes(jc, oracle, name, lattice):
    """Device-side VolumeParticleEmitter3 over analytic surfaces (jetcu_hybrid_emit_shapes): same points, same order."""
    n = 16
    prm = HybridParams.make(Grid.cube(n), mode=PIC, mgpcg=MgpcgParams.make(max_levels=2))
    a, b = jc.hybrid(prm), oracle.hybrid(prm)
    na, nb = (s.emit_shapes(EMITTERS[name], 0.5 / n, (0.1, 0.0, -0.2), lattice=lattice) for s in (a, b))
    (pa, va), (pb, vb) = a.get_particles(), b.get_particles()
    if any(m["kind"] == "cylinder" for m in EMITTERS[name]):
        gc.same_points_up_to_cylinder_surface(pa, pb, EMITTERS[name])
    else:
        assert na == nb
        exact(pa, pb, "emitted positions"); exact(va, vb, "emitted velocities")
    a.close(); b.close()


def test_save_positions(jc, tmp_path):
    """.pos / .xyz frames from the device-resident set == the frame of the positions get_particles returns."""
    n = 16
    h = jc.hybrid(HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_levels=3)))
    h.emit_shapes(EMITTERS["sphere"], 0.5 / n)
    h.step(1.0 / 60.0)
    pos, _ = h.get_particles()
    h.save_positions(str(tmp_path / "f.pos"), "pos")
    h.save_positions(str(tmp_path / "f.xyz"), "xyz")
    raw = (tmp_path / "f.pos").read_bytes()
    assert raw[:20].hex() == "0c00000000000600080004000600000004000000"
    assert int.from_bytes(raw[20:24], "little") == pos.size * 8 and raw[24:] == pos.tobytes()
    text = (tmp_path / "f.xyz").read_text().splitlines()
    assert len(text) == len(pos) and text[0] == "%g %g %g" % tuple(pos[0]) and text[-1] == "%g %g %g" % tuple(pos[-1])
    h.close()


def test_golden_emission(jc):
    gc.emission(jc.hybrid, exact_bits=False)

