# BASELINE configs[0]: FLIP dam break at 64^3, h/2 lattice (8 particles per cell), 60 frames at 60 fps with the reference's
# adaptive sub-stepping: the C restatement (oracle/libjetport.so) against the unmodified reference (oracle/_ref), frame by frame,
# bit for bit.  CPU only, ~5 minutes:   python scripts/config0_port_vs_reference.py 60 > profiles/r01_config0_port_vs_reference.log
import sys, time; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np
from oracle_api import *
ref, port = load_ref(), load_port()
ref.set_num_threads(0); port.set_num_threads(0)
n = 64
prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=5))
a, b = ref.hybrid(prm), port.hybrid(prm)
na = a.emit_box((0,0,0),(0.4,0.6,1.0),0.5/n); nb = b.emit_box((0,0,0),(0.4,0.6,1.0),0.5/n)
print("particles", na, nb, flush=True)
t0=time.time(); tot=0
for frame in range(int(sys.argv[1]) if len(sys.argv)>1 else 60):
    ka, sa = a.advance_frame(1/60.); kb, sb = b.advance_frame(1/60.)
    tot += kb
    pa, va = a.get_particles(); pb, vb = b.get_particles()
    same = np.array_equal(pa,pb) and np.array_equal(va,vb) and ka==kb and sa.pcg_iterations==sb.pcg_iterations
    print("frame %d: substeps %d/%d iters %d/%d %s  (%.0f s)" % (frame, ka, kb, sa.pcg_iterations, sb.pcg_iterations, "IDENTICAL" if same else "DIFFERENT max|dx| %.3e" % np.abs(pa-pb).max(), time.time()-t0), flush=True)
    if not same: break
print("total sub-steps", tot)
