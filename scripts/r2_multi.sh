#!/bin/bash
# Round 2 multi-GPU call: gpurun --gpus N -- 'bash scripts/r2_multi.sh N [what...]'
#   what: parity latency bench256 bench512 smoke (default: parity latency bench256)
set -u
N=${1:-2}; shift || true
WHAT=${*:-parity latency bench256}
OUT=gpurun_out/r2_multi_$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-600)"; }
nvidia-smi -L | head -8
for w in $WHAT; do case $w in
parity)
  run pytest_slabs 900 python -m pytest tests/test_gpu_slabs.py -m gpu -q -x --tb=short -p no:cacheprovider ;;
latency)
  for p in 0 1; do JETCU_P2P_HALO=$p run halo_latency_p2p$p 200 $TR --master-port 2953$p scripts/halo_latency.py; done ;;
bench256)
  run bench256 200 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 ;;
bench256_p2p)
  JETCU_P2P_HALO=1 run bench256_p2p 200 $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 ;;
knobs)
  JETCU_SLAB_MIN_CELLS=3000000 run bench256_replicate128 200 $TR --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 0
  JETCU_RELAX_BLOCK_DIV=4 run bench256_blockdiv4 200 $TR --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 0 ;;
bench512)
  run bench512 420 $TR --master-port 29545 bench.py --gpus $N --grid 512 --steps 5 --warmup 2 --e2e-steps ${E2E512:-1} ;;
smoke)
  S=${SMOKE_N:-768}
  run smoke$S 420 $TR --master-port 29546 scripts/bench_smoke.py $S 4 2 ;;
esac; done
python - $OUT <<'PY'
import json, sys, glob, os
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.log"))):
    for l in open(f, errors="replace"):
        l = l.strip()
        if l.startswith("{") and '"metric"' in l:
            d = json.loads(l)
            print(os.path.basename(f), d["n_gpus"], "GPUs", round(d["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in d.get("stages_ms", {}).items() if isinstance(v, float)},
                  d.get("parity_check"), (d.get("comm") or {}))
PY
ls -la $OUT
