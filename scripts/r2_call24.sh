#!/bin/bash
# Round 2, GPU call 24 (one B200): the interior / bulk fast path of the SOR cell update (relax_cell_code) — parity and timing.
set -u
OUT=gpurun_out/r2c25
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; local t0=$SECONDS; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $((SECONDS-t0))s $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_linear 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
ab fast X=0




python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s n %4d total %8.2f press %8.2f solve %8.2f iters %s" % (d["tag"], d["n"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"]))
PY
tail -3 $OUT/ab.err
