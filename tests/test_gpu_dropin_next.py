"""Jet-side drop-in check of the adapter paths added after the three seams (adapters/dropin_check_next.cpp, public Jet API
only): a FLIP dam break onto a RigidBodyCollider3(Box3) through jetcu::CudaFlipSolver3, and jet::GridSmokeSolver3 vs
jetcu::CudaGridSmokeSolver3, each compared with the stock CPU solver in the same process.
"""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "adapters", "_build", "dropin_check_next")
pytestmark = pytest.mark.gpu


def run(*args):
    if not os.path.exists(BIN):
        pytest.fail("adapters/_build/dropin_check_next missing: run `make -C adapters` where /root/reference exists")
    r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_collider_scene_matches_stock_jet():
    out = run("--what", "collider", "--n", 32, "--frames", 3, "--levels", 4)
    assert out["particles"] > 10000
    assert out["max_dx"] <= 1e-5, out          # BASELINE tolerance: 1e-5 of the domain size (= 1)


@pytest.mark.parametrize("what", ["pillars", "tank"])
def test_example_colliders_match_stock_jet(what):
    """The colliders of hybrid_liquid_sim examples 3 / 6 (ImplicitSurfaceSet3 of Cylinder3; normal-flipped Sphere3)."""
    out = run("--what", what, "--n", 32, "--frames", 3, "--levels", 4)
    assert out["particles"] > 10000
    assert out["max_dx"] <= 1e-5, out


def test_device_emission_matches_host_emitters():
    """Example 1 / 2's emitter set (half-space + sphere) on the host vs CudaFlipSolver3::emitOnDevice, then 3 frames."""
    out = run("--what", "emit", "--n", 32, "--frames", 3, "--levels", 4)
    assert out["particles"] > 10000
    assert out["max_dx"] <= 1e-5, out


@pytest.mark.parametrize("what", ["smoke", "smoke_collider", "smoke_device"])
def test_smoke_matches_stock_jet(what):
    out = run("--what", what, "--n", 32, "--frames", 5, "--levels", 4)
    assert out["velocity_scale"] > 0.05        # the plume really rose
    assert out["velocity_rel"] <= 1e-9 and out["density_rel"] <= 1e-9, out


APP = os.path.join(ROOT, "adapters", "_build", "hybrid_liquid_sim_cuda")


@pytest.mark.parametrize("example", [1, 3, 6])
def test_example_application_next_to_stock_jet(example, tmp_path):
    """adapters/hybrid_liquid_sim_cuda: the scenes of Jet's hybrid_liquid_sim example on jetcu's solver classes — device
    emission, device collider, frames written from the device — stepped next to a stock Jet solver (--compare)."""
    if not os.path.exists(APP):
        pytest.fail("adapters/_build/hybrid_liquid_sim_cuda missing: run `make -C adapters` where /root/reference exists")
    r = subprocess.run([APP, "-e", str(example), "-r", "32", "-f", "3", "-o", str(tmp_path), "-m", "pos", "--compare"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    frames = [l for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert len(frames) == 3, r.stdout
    for l in frames:
        dx = float(l.rsplit(" ", 1)[1])
        assert 0.0 <= dx <= 1e-5, l                     # -1 = particle counts differ
    first = (tmp_path / "frame_000000.pos").read_bytes()
    assert first[:20].hex() == "0c00000000000600080004000600000004000000" and (len(first) - 24) % 24 == 0 and len(first) > 24 * 1000


def test_example_application_with_its_own_solver_configuration(tmp_path):
    """hybrid_liquid_sim example 3 exactly as Jet ships it (main.cpp:219): a default-constructed pressure solver, i.e.
    FdmIccgSolver3(100, 1e-6), and setUseCompressedLinearSystem(true) on both sides — stock Jet builds the compressed CSR system,
    the adapter forwards ICCG + the flag to the device."""
    r = subprocess.run([APP, "-e", "3", "-r", "32", "-f", "3", "-o", str(tmp_path), "-m", "none", "--compare", "--iccg"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    frames = [l for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert len(frames) == 3, r.stdout
    for l in frames:
        assert 0.0 <= float(l.rsplit(" ", 1)[1]) <= 1e-5, l


@pytest.mark.parametrize("flags", [[], ["--linear"]])
def test_smoke_application_next_to_stock_jet(flags):
    """adapters/smoke_sim_cuda: smoke_sim example 1 (continuous box source under a sphere) with the source on the device, stepped
    next to a stock GridSmokeSolver3 with Jet's own VolumeGridEmitter3."""
    app = os.path.join(ROOT, "adapters", "_build", "smoke_sim_cuda")
    if not os.path.exists(app):
        pytest.fail("adapters/_build/smoke_sim_cuda missing: run `make -C adapters` where /root/reference exists")
    r = subprocess.run([app, "-r", "32", "-f", "6", "--compare"] + flags, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    frames = [l for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert len(frames) == 6, r.stdout
    for l in frames:
        assert float(l.rsplit(" ", 1)[1]) <= 1e-8, l

