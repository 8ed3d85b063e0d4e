#!/bin/bash
# Round 2 multi-GPU call C: gpurun --gpus N -- 'bash scripts/r2_multi_c.sh N' — slab parity tests and the contract bench on the
# final build (dependent launches, device-side alpha) at N ranks, where the middle ranks have two neighbours.
set -u
N=${1:-4}
OUT=gpurun_out/r2_multic_$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name=$1; shift; echo "== $name"; local t0=$SECONDS; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $((SECONDS-t0))s $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run pytest_slabs 400 python -m pytest tests/test_gpu_slabs.py -m gpu -q -x --tb=short -p no:cacheprovider
run bench256 200 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-projection
python - $OUT <<'PY'
import json, sys, glob, os
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench*.log"))):
    for l in open(f, errors="replace"):
        l = l.strip()
        if l.startswith("{") and '"metric"' in l:
            d = json.loads(l)
            print(os.path.basename(f), d["n_gpus"], "GPUs", round(d["ms_per_step"], 2), "ms/step", {k: round(v, 2) for k, v in d.get("stages_ms", {}).items() if isinstance(v, float)},
                  d.get("parity_check"), d.get("fixed_work"), (d.get("comm") or {}))
PY
