#!/bin/bash
# Round 2 multi-GPU call B: gpurun --gpus N -- 'bash scripts/r2_multi_b.sh N'
# slab parity on the current build, bench at 256^3, and the binning sub-stage trace (JETCU_BIN_TRACE=1, synchronising).
set -u
N=${1:-2}
OUT=gpurun_out/r2_multib_$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-500)"; }
run pytest_slabs 600 python -m pytest tests/test_gpu_slabs.py -m gpu -q -x --tb=short -p no:cacheprovider
run bench256 200 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-projection
JETCU_BIN_TRACE=1 run bench256_trace 200 $TR --master-port 29542 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-projection --e2e-steps 0
grep "bin trace rank 0" $OUT/bench256_trace.log | tail -8
python - $OUT <<'PY'
import json, sys, glob, os
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench*.log"))):
    for l in open(f, errors="replace"):
        l = l.strip()
        if l.startswith("{") and '"metric"' in l:
            d = json.loads(l)
            print(os.path.basename(f), d["n_gpus"], "GPUs", round(d["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in d.get("stages_ms", {}).items() if isinstance(v, float)},
                  d.get("parity_check"), (d.get("comm") or {}))
PY
