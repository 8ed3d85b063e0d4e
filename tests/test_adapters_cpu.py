"""CPU-side facts the adapters rely on, checked on the UNMODIFIED reference through the Jet-side check binary (no GPU):
with a multigrid linear-system solver installed, GridFractionalSinglePhasePressureSolver3 ignores useCompressed
(src/jet/grid_fractional_single_phase_pressure_solver3.cpp:465-481, 651-690) — which is why
jetcu::CudaGridFractionalSinglePhasePressureSolver3 may accept the flag Jet's examples set."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "adapters", "_build", "dropin_check_next")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="adapters/_build/dropin_check_next not built (needs /root/reference)")


def test_reference_ignores_use_compressed_with_mgpcg():
    r = subprocess.run([BIN, "--what", "compressed", "--n", "24", "--frames", "4", "--levels", "3"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["particles"] > 10000 and out["different"] == 0, out
