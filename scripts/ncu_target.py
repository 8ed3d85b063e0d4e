#!/usr/bin/env python
"""Small ncu target: one config-2 projection (flat slab + box collider) at n^3 through jetcu_pressure_solve with
`iters` PCG iterations — the first k_relax_colour_code launches are the finest-level colour passes.
    ncu --set full -k regex:k_relax_colour_code -c 2 ... python scripts/ncu_target.py 256 1"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jetcu_loader import load_package

load_package()
from fluid_engine_dev_b200 import Jetcu
from fluid_engine_dev_b200.jetcu import MgpcgParams
import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
jc = Jetcu(0)
sc = scenes.projection_scene(n, "slab", True)
levels = 1
m = n
while levels < 7 and m % 2 == 0 and m > 4:
    m //= 2
    levels += 1
prm = MgpcgParams.make(max_iterations=iters, max_levels=levels)
_, _, _, _, res = jc.pressure_solve(sc["grid"], sc["u"], sc["v"], sc["w"], prm, collider_sdf=sc["collider_sdf"],
                                    fluid_sdf=sc["fluid_sdf"])
print(res.as_dict())
