#!/bin/bash
# Multi-GPU call of round 2 (run with `gpurun --gpus 2 --timeout 1200 -- 'bash scripts/round2_multi_gpu_call.sh 2'`, later with 8):
# what session 9 prepared for the slab path without a GPU.  Every step has its own timeout; results land in gpurun_out/r2_multi/.
set -u
N=${1:-2}
OUT=gpurun_out/r2_multi_$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -1 $OUT/$name.log | cut -c1-400)"; }

# 1. parity: slabs == single GPU, with NCCL halos and with the peer-memory halos (first run of the latter: non-strict xfail)
run pytest_slabs 700 python -m pytest tests/test_gpu_slabs.py -m gpu -q -rxX
# 2. per-exchange latency of both transports
for p in 0 1; do JETCU_P2P_HALO=$p run halo_latency_p2p$p 200 $TR --master-port 2953$p scripts/halo_latency.py; done
# 3. bench: default plan (MG levels of 64^3 and below replicated) with NCCL halos, with peer-memory halos, and the round-1 plan
run bench_default 300 $TR --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3
JETCU_P2P_HALO=1 run bench_p2p 300 $TR --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3
JETCU_SLAB_MIN_CELLS=8192 run bench_round1_plan 300 $TR --master-port 29543 bench.py --gpus $N --steps 5 --warmup 3
JETCU_SLAB_MIN_CELLS=3000000 run bench_replicate_128 300 $TR --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3
grep -h '"metric"' $OUT/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], 'GPUs', round(d['ms_per_step'], 2), 'ms/step  pressure', round(d['stages_ms']['ms_pressure'], 2), d.get('comm', {}).get('backend', '')[:40])
"
