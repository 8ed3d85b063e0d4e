"""Golden for the BENCHMARK-SIZE workload: bench.py's FLIP 256^3 tank (126 584 289 particles, at rest) stepped by the
UNMODIFIED reference (oracle/_ref/libjetref.so) for the first K sub-steps.  Run here (the container with /root/reference):

    make -C oracle ref && python tests/golden/make_golden_256.py [K=3] [mode=flip|apic] [v0x=0.0]

A 256^3 step of the reference costs minutes and 30+ GB of host memory, so the GPU box never runs it inside a test: the
committed fixture carries, per sub-step, the PCG iteration count and residuals, every 65 521-th particle (position and
velocity, caller's order), a 16^3 sub-lattice of the pressure field, the per-axis position sums and the wall-clock
seconds of the step (the latter feed nothing but profiles/).  tests/test_gpu_fullsize.py compares the CUDA path with it."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle_api import APIC, FLIP, Grid, HybridParams, MgpcgParams, load_ref  # noqa: E402

N = 256
STRIDE = 65521
PSTRIDE = 16


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    mode_name = sys.argv[2] if len(sys.argv) > 2 else "flip"
    mode = APIC if mode_name == "apic" else FLIP
    v0x = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    ref = load_ref()
    assert ref.kind == "reference"
    ref.set_num_threads(0)
    prm = HybridParams.make(Grid.cube(N), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=6))
    h = ref.hybrid(prm)
    t0 = time.perf_counter()
    n = h.emit_box((0.0, 0.0, 0.0), (1.0, 0.9375, 1.0), 0.5 / N, (v0x, 0.0, 0.0))
    print("emitted %d particles in %.1f s" % (n, time.perf_counter() - t0), flush=True)
    idx = np.arange(0, n, STRIDE)
    out = dict(n=np.array(n), idx=idx, dt=np.array(1.0 / 240.0), threads=np.array(ref.num_threads()), v0x=np.array(v0x))
    log = []
    for s in range(k):
        t0 = time.perf_counter()
        st = h.step(1.0 / 240.0)
        sec = time.perf_counter() - t0
        pos, vel = h.get_particles()
        pr = h.get_pressure()
        d = st.as_dict()
        out["pos%d" % s] = pos[idx].copy()
        out["vel%d" % s] = vel[idx].copy()
        out["possum%d" % s] = pos.sum(axis=0, dtype=np.float64)
        out["pressure%d" % s] = pr[PSTRIDE // 2::PSTRIDE, PSTRIDE // 2::PSTRIDE, PSTRIDE // 2::PSTRIDE].copy()
        out["pmax%d" % s] = np.array(np.abs(pr).max())
        out["iters%d" % s] = np.array(int(d["pcg_iterations"]))
        out["last_residual%d" % s] = np.array(d["pcg_last_residual"])
        out["true_residual%d" % s] = np.array(d["pcg_true_residual_l2"])
        out["b_l2%d" % s] = np.array(d["pcg_b_l2"])
        out["seconds%d" % s] = np.array(sec)
        log.append(dict(step=s, seconds=sec, **{kk: (float(v) if isinstance(v, float) else int(v)) for kk, v in d.items()}))
        print(json.dumps(log[-1]), flush=True)
        del pos, vel, pr
        np.savez_compressed(os.path.join(HERE, "fullsize_%s_256%s.npz" % (mode_name, "" if v0x == 0.0 else "_slosh")), steps=np.array(s + 1), **out)
    h.close()


if __name__ == "__main__":
    main()
