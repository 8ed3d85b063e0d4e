"""CPU checks of two claims the CUDA path relies on, made on the ORACLE's own outputs, plus the reference arm of bench.py.

1. Bulk-diagonal shortcut of the SOR kernel (k_relax_colour_code, csrc/fdm.cu): a cell whose six faces are fully open
   liquid-liquid couplings has the per-level constant diagonal ((((((0+sx)+sx)+sy)+sy)+sz)+sz).  Checked on the matrices
   GridFractionalSinglePhasePressureSolver3::buildSingleSystem produces (through the oracle) on every MG level of several
   scenes: wherever all six couplings equal -invH^2 exactly, the stored center must equal that constant bit for bit.
2. The face-code classes the packed matrix uses are exhaustive: every off-diagonal of those matrices is either 0 or
   exactly 0.0 - double(float weight) * invH^2.
3. `bench.py --impl reference` prints one well-formed contract line (bounded CPU sample)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import scenes
from oracle_api import MgpcgParams, have_port, load_port

pytestmark = pytest.mark.skipif(not have_port(), reason="oracle/libjetport.so not built")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def port():
    o = load_port()
    o.set_num_threads(0)
    return o


@pytest.mark.parametrize("n,kind,collider,domain", [(32, "slab", True, 1.0), (32, "column", True, 1.0), (24, "slab", False, 0.9),
                                                    (16, "column", False, 1.0)])
def test_bulk_diagonal_is_the_level_constant(port, n, kind, collider, domain):
    sc = scenes.projection_scene(n, kind, collider)
    grid = sc["grid"]
    if domain != 1.0:   # a non-power-of-two spacing: the constant is then not a round number
        grid.hx = grid.hy = grid.hz = domain / n
    levels = port.mg_num_levels(n, n, n, 4)
    port.pressure_solve(grid, sc["u"], sc["v"], sc["w"], MgpcgParams.make(max_iterations=2, max_levels=levels),
                        collider_sdf=sc["collider_sdf"], fluid_sdf=sc["fluid_sdf"])
    h = grid.hx
    checked = 0
    for l in range(levels):
        sysl = port.pressure_get_system(grid, l)
        A = sysl["A"]
        inv = 1.0 / h
        s = inv * inv                      # invHSqr = invH * invH (…pressure_solver3.cpp:127-128)
        c6 = np.float64(0.0)
        for _ in range(6):
            c6 = c6 + np.float64(s)        # +x, -x, +y, -y, +z, -z, every term = 1.0f * invH^2
        r, u, f, c = A["right"], A["up"], A["front"], A["center"]
        full = np.zeros(c.shape, dtype=bool)
        full[1:-1, 1:-1, 1:-1] = ((r[1:-1, 1:-1, 1:-1] == -s) & (r[1:-1, 1:-1, :-2] == -s) & (u[1:-1, 1:-1, 1:-1] == -s) &
                                  (u[1:-1, :-2, 1:-1] == -s) & (f[1:-1, 1:-1, 1:-1] == -s) & (f[:-2, 1:-1, 1:-1] == -s))
        if full.any():
            assert np.all(c[full] == c6), "level %d: %d bulk cells, center range %r vs %r" % (l, full.sum(), (c[full].min(), c[full].max()), c6)
            checked += int(full.sum())
        # every off-diagonal is 0 or exactly -(double(float w) * invH^2) for the float face weight of that face
        uw, vw, ww = sysl["u_weights"], sysl["v_weights"], sysl["w_weights"]
        nx = c.shape[2]
        exp_r = np.where(r != 0, -(uw[:, :, 1:nx + 1].astype(np.float64) * s), 0.0)
        exp_u = np.where(u != 0, -(vw[:, 1:, :].astype(np.float64) * s), 0.0)
        exp_f = np.where(f != 0, -(ww[1:, :, :].astype(np.float64) * s), 0.0)
        assert np.array_equal(r, exp_r) and np.array_equal(u, exp_u) and np.array_equal(f, exp_f), "level %d" % l
        h *= 2.0
    assert checked > 100


def test_bench_reference_arm_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--sample-n", "16"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
