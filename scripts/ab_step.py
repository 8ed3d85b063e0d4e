#!/usr/bin/env python
"""A/B stage timings of the bench scenes without torch: emits the scene, W warm-up sub-steps, K timed ones, prints the mean
stage times of the library's own CUDA-event timers as one JSON line.  Environment knobs (JETCU_*) are read by the library
once per process, so every configuration is its own process.
    JETCU_MOVE_CACHE=0 python scripts/ab_step.py --n 256 --mode flip --workload tank --steps 6 --warmup 3 --tag base"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jetcu_loader import load_package

load_package()
from fluid_engine_dev_b200 import Jetcu
from fluid_engine_dev_b200.jetcu import FLIP, APIC, PIC, Grid, HybridParams, MgpcgParams

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--mode", default="flip")
ap.add_argument("--workload", default="tank")
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--dt", type=float, default=1.0 / 240.0)
ap.add_argument("--tag", default="")
ap.add_argument("--iccg", action="store_true", help="FdmIccgSolver3(100, 1e-6) + compressed flag: what Jet's hybrid_liquid_sim examples run")
ap.add_argument("--natural-ordering", action="store_true", help="useRedBlackOrdering = false (FdmMgpcgSolver3's constructor default)")
a = ap.parse_args()
scenes = {"tank": ((0.0, 0.0, 0.0), (1.0, 0.9375, 1.0), (0.0, 0.0, 0.0)),
          "slosh": ((0.0, 0.0, 0.0), (1.0, 0.9375, 1.0), (0.5, 0.0, 0.0)),
          "dam": ((0.0, 0.0, 0.0), (0.4, 0.6, 1.0), (0.0, 0.0, 0.0))}
lo, hi, v0 = scenes[a.workload]
jc = Jetcu(0)
levels, m = 1, a.n
while levels < 6 and m % 2 == 0 and m > 4:
    m //= 2
    levels += 1
mode = {"flip": FLIP, "apic": APIC, "pic": PIC}[a.mode]
solver = MgpcgParams.iccg(100, 1e-6, compressed=True) if a.iccg else MgpcgParams.make(max_iterations=100, max_levels=levels, use_red_black=not a.natural_ordering)
h = jc.hybrid(HybridParams.make(Grid.cube(a.n), mode=mode, mgpcg=solver))
npart = h.emit_box(lo, hi, 0.5 / a.n, v0)
for _ in range(a.warmup):
    h.step(a.dt)
st = [h.step(a.dt).as_dict() for _ in range(a.steps)]
keys = ("ms_total", "ms_bin", "ms_p2g", "ms_sdf", "ms_extrapolate", "ms_pressure", "ms_move", "ms_mgpcg_solve")
out = {"tag": a.tag, "n": a.n, "mode": a.mode, "workload": a.workload, "particles": int(npart),
       "env": {k: v for k, v in os.environ.items() if k.startswith("JETCU_")},
       "iters": [int(s["pcg_iterations"]) for s in st]}
for k in keys:
    out[k] = round(sum(s[k] for s in st) / len(st), 3)
if os.environ.get("AB_HASH"):   # bit-identity checks between builds / switches: digests of the final state
    import hashlib
    import numpy as np
    pos, vel = h.get_particles()
    out["sha_pos"] = hashlib.sha1(np.ascontiguousarray(pos).tobytes()).hexdigest()
    out["sha_vel"] = hashlib.sha1(np.ascontiguousarray(vel).tobytes()).hexdigest()
    out["sha_pressure"] = hashlib.sha1(np.ascontiguousarray(h.get_pressure()).tobytes()).hexdigest()
print(json.dumps(out))
h.close()
