#!/bin/bash
# Round 2, GPU call 22 (one B200): APIC 256^3 against the reference golden (BASELINE config 3); the row-per-thread natural-ordering
# tile kernel (parity + timing against the 512-thread tile kernel, JETCU_LEX_ROWS=0).
set -u
OUT=gpurun_out/r2c22
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; local t0=$SECONDS; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $((SECONDS-t0))s $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_apic_golden 400 python -m pytest "tests/test_gpu_fullsize.py::test_flip_256_matches_reference_golden" -m gpu -q -x --tb=short -p no:cacheprovider -k apic
run pytest_lex_rows 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py tests/test_gpu_dropin.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 128 --mode flip --workload tank --steps 4 --warmup 2 --natural-ordering"
ab lex128_rows X=0
ab lex128_tiles JETCU_LEX_ROWS=0
ABARGS="--n 256 --mode flip --workload tank --steps 3 --warmup 1 --natural-ordering"
ab lex256_rows X=0
ab lex256_tiles JETCU_LEX_ROWS=0
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s n %4d total %8.2f press %8.2f solve %8.2f iters %s" % (d["tag"], d["n"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"]))
PY
tail -3 $OUT/ab.err
