#!/usr/bin/env python
"""ncu target: the bench scene (FLIP tank) at n^3, `steps` sub-steps through jetcu_hybrid_step.
    ncu --set full --clock-control none --import-source on -k regex:k_p2g_cells -c 1 -o gpurun_out/p2g python scripts/ncu_step.py 256 1
Graphs are off (JETCU_NO_GRAPHS=1) so that every kernel is a plain launch ncu can filter by name."""
import os
import sys

os.environ.setdefault("JETCU_NO_GRAPHS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jetcu_loader import load_package

load_package()
from fluid_engine_dev_b200 import Jetcu
from fluid_engine_dev_b200.jetcu import FLIP, APIC, Grid, HybridParams, MgpcgParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode = APIC if (len(sys.argv) > 3 and sys.argv[3] == "apic") else FLIP
solver_kind = sys.argv[4] if len(sys.argv) > 4 else "mgpcg"   # mgpcg | natural (useRedBlackOrdering = false) | iccg
jc = Jetcu(0)
levels, m = 1, n
while levels < 6 and m % 2 == 0 and m > 4:
    m //= 2
    levels += 1
solver = (MgpcgParams.iccg(100, 1e-6, compressed=True) if solver_kind == "iccg"
          else MgpcgParams.make(max_iterations=100, max_levels=levels, use_red_black=solver_kind != "natural"))
h = jc.hybrid(HybridParams.make(Grid.cube(n), mode=mode, mgpcg=solver))
npart = h.emit_box((0.0, 0.0, 0.0), (1.0, 0.9375, 1.0), 0.5 / n, (0.0, 0.0, 0.0))
for s in range(steps):
    st = h.step(1.0 / 240.0)
    print(npart, st.as_dict())
