"""GPU drop-in check: UNMODIFIED Jet host code (adapters/dropin_check.cpp, public Jet API only, linked against the
reference compiled from /root/reference) drives the jetcu library through the three C++ adapter seams of
adapters/jetcu_jet_adapters.h and is compared with the stock CPU solver in the same process.

Tolerances are BASELINE.json's: pressure <= 1e-6 relative L-inf, particle positions <= 1e-5 of the domain size (= 1)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "adapters", "_build", "dropin_check")


def run(*args):
    if not os.path.exists(BIN):
        pytest.fail("adapters/_build/dropin_check missing: run `make -C adapters` where /root/reference exists "
                    "(__graft_entry__.build() does)")
    r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["flip", "apic", "pic"])
def test_three_seams_match_stock_jet(mode):
    out = run("--n", 32, "--frames", 3, "--levels", 4, "--mode", mode)
    assert out["particles"] == 65910
    for seam in ("seam1", "seam2", "seam3"):
        assert out[seam + "_max_dx"] <= 1e-5, out
        assert out[seam + "_last_iterations"] == out["stock_last_iterations"], out
    for seam in ("seam1", "seam2"):
        assert out[seam + "_pressure_rel"] <= 1e-6, out


@pytest.mark.gpu
def test_three_seams_natural_ordering():
    """FdmMgpcgSolver3's constructor default, useRedBlackOrdering = false (include/jet/fdm_mgpcg_solver3.h:41): the serial
    FdmGaussSeidelSolver3::relax on the CPU, the tile-hyperplane wavefront on the device."""
    out = run("--n", 32, "--frames", 2, "--levels", 4, "--mode", "flip", "--natural-ordering")
    for seam in ("seam1", "seam2", "seam3"):
        assert out[seam + "_max_dx"] <= 1e-5, out
        assert out[seam + "_last_iterations"] == out["stock_last_iterations"], out
    for seam in ("seam1", "seam2"):
        assert out[seam + "_pressure_rel"] <= 1e-6, out


@pytest.mark.gpu
def test_flip_64_ten_frames_device_resident():
    out = run("--n", 64, "--frames", 10, "--levels", 5, "--mode", "flip", "--skip-seam12")
    assert out["particles"] == 516516
    assert out["seam3_max_dx"] <= 1e-5, out


def test_no_cpu_fallback():
    """Without a usable GPU the adapters must fail loudly (exit code 1, 'no CPU fallback'), never compute on the host."""
    if not os.path.exists(BIN):
        pytest.skip("drop-in binary not built here")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([BIN, "--n", "8", "--frames", "1", "--levels", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr
