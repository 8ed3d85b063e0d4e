#!/usr/bin/env python
"""BASELINE config 5 probe: device-resident GridSmokeSolver3 at n^3 (default 256), plume rising from rest in a closed box.
    python scripts/bench_smoke.py [n] [steps] [warmup] [plume|example1]
`example1` = src/examples/smoke_sim example 1: a continuous box source (device-side emitter, fired before every step) under a
sphere collider, starting from empty fields.
Prints one JSON object: ms/step (CUDA events on the library stream), Mcell/s, PCG iterations per step.  Not the contract
bench (bench.py measures BASELINE's FLIP metric); a measuring stick for the smoke kernels and an ncu target."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from jetcu_loader import load_package

load_package()
from fluid_engine_dev_b200 import Jetcu
from fluid_engine_dev_b200.jetcu import Grid, MgpcgParams, SmokeParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scene = sys.argv[4] if len(sys.argv) > 4 else "plume"
levels, m = 1, n
while levels < 6 and m % 2 == 0 and m > 4:
    m //= 2
    levels += 1
jc = Jetcu(0)
prm = SmokeParams.make(Grid.cube(n), mgpcg=MgpcgParams.make(max_iterations=100, max_levels=levels))
s = jc.smoke(prm)
c = (np.arange(n) + 0.5) / n
Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
den = np.ascontiguousarray(np.clip(1.0 - np.sqrt((X - 0.5) ** 2 + (Y - 0.25) ** 2 + (Z - 0.45) ** 2) / 0.2, 0.0, 1.0))
source = [dict(kind="box", a=(0.45, -1.0, 0.45), b=(0.55, 0.05, 0.55))]
if scene == "example1":
    s.set_collider_shapes([dict(kind="sphere", a=(0.5, 0.3, 0.5), b=(0.075, 0.0, 0.0))])
else:
    s.set_fields(den, 2.0 * den + 0.1)
dt = 1.0 / 60.0


def step():
    if scene == "example1":
        s.emit_step(source, (0.0, 1.0), (0.0, 1.0))
    return s.step(dt)


for _ in range(warmup):
    step()
stream = torch.cuda.ExternalStream(jc.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
l0 = jc.launch_count()
e0.record(stream)
iters = [int(step().iterations) for _ in range(steps)]
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"workload": "GridSmokeSolver3 %d^3, %s, closed box, MGPCG(100, %d levels)" %
                              (n, "smoke_sim example 1 (box source under a sphere collider)" if scene == "example1" else "plume from rest", levels),
                  "ms_per_step": ms, "mcell_per_s": n ** 3 / ms / 1e3, "pcg_iterations": iters,
                  "gpu_launches": jc.launch_count() - l0}))
