#!/bin/bash
# Round 2, GPU call 19 (one B200): shared-memory staged natural-ordering tiles and single-CTA coarse-level smoothers.
set -u
OUT=gpurun_out/r2c19
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_linear 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py tests/test_gpu_dropin.py tests/test_gpu_smoke.py tests/test_gpu_diffusion.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
ab base X=0
ABARGS="--n 64 --mode flip --workload dam --steps 6 --warmup 3"
ab dam64 X=0
ABARGS="--n 128 --mode flip --workload tank --steps 4 --warmup 2 --natural-ordering"
ab lex128 X=0
ABARGS="--n 256 --mode flip --workload tank --steps 3 --warmup 1 --natural-ordering"
ab lex256 X=0
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s n %4d total %8.2f press %8.2f solve %8.2f iters %s" % (d["tag"], d["n"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"]))
PY
tail -3 $OUT/ab.err
