"""Which 256^3 scenes does the faithful MGPCG converge on?  (GPU only; prints iterations and true residuals.)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from jetcu_loader import load_package
load_package()
from fluid_engine_dev_b200 import Jetcu
from fluid_engine_dev_b200.jetcu import FLIP, Grid, HybridParams, MgpcgParams
jc = Jetcu(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cases = [
    ("tank .9375 v0=.5 L6", (1, .9375, 1), (.5, 0, 0), 6, 1/240),
    ("tank .9375 v0=0 L6", (1, .9375, 1), (0, 0, 0), 6, 1/240),
    ("tank .5 v0=0 L6", (1, .5, 1), (0, 0, 0), 6, 1/240),
    ("tank .5 v0=0 L7", (1, .5, 1), (0, 0, 0), 7, 1/240),
    ("tank .5 v0=.5 L6", (1, .5, 1), (.5, 0, 0), 6, 1/240),
    ("dam L6", (.4, .6, 1), (0, 0, 0), 6, 1/240),
    ("dam L5", (.4, .6, 1), (0, 0, 0), 5, 1/240),
    ("dam L7", (.4, .6, 1), (0, 0, 0), 7, 1/240),
    ("tank .9375 v0=0 L5", (1, .9375, 1), (0, 0, 0), 5, 1/240),
    ("tank .9375 v0=0 L4", (1, .9375, 1), (0, 0, 0), 4, 1/240),
    ("tank .93 v0=0 L6", (1, .93, 1), (0, 0, 0), 6, 1/240),
    ("tank .93 v0=.5 L6", (1, .93, 1), (.5, 0, 0), 6, 1/240),
    ("tank .9 v0=.5 L6", (1, .9, 1), (.5, 0, 0), 6, 1/240),
    ("tank .75 v0=.5 L6", (1, .75, 1), (.5, 0, 0), 6, 1/240),
    ("tank .9375 v0=.5 L6 dt/2", (1, .9375, 1), (.5, 0, 0), 6, 1/480),
    ("tank .9375 v0=.5 L6 dt*2", (1, .9375, 1), (.5, 0, 0), 6, 1/120),
    ("tank .9375 v0=.5 L7", (1, .9375, 1), (.5, 0, 0), 7, 1/240),
    ("tank .9375 v0=(.5,.3,.2) L6", (1, .9375, 1), (.5, .3, .2), 6, 1/240),
    ("full 1.0 v0=.5 L6", (1, 1, 1), (.5, 0, 0), 6, 1/240),
]
if len(sys.argv) > 2:
    cases = [c for c in cases if sys.argv[2] in c[0]]
for name, up, v0, L, dt in cases:
    prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=L))
    h = jc.hybrid(prm)
    npart = h.emit_box((0, 0, 0), up, 0.5 / n, v0)
    out = []
    for s in range(8):
        st = h.step(dt)
        out.append("%d it res %.1e true %.1e (%.0f ms, press %.0f)" % (st.pcg_iterations, st.pcg_last_residual, st.pcg_true_residual_l2 / max(st.pcg_b_l2, 1e-300), st.ms_total, st.ms_pressure))
    print(name, npart, " | ".join(out), flush=True)
    h.close()
