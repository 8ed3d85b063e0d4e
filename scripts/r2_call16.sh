#!/bin/bash
# Round 2, GPU call 16 (one B200): the natural (lexicographic) Gauss-Seidel ordering on the device — parity tests (relax, V-cycle,
# MGPCG, drop-in seams with FdmMgpcgSolver3's constructor default), memcheck of the wavefront kernels, timing at 128^3 / 256^3.
set -u
OUT=gpurun_out/r2c16
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_linear 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py tests/test_gpu_dropin.py -m gpu -q -x --tb=short -p no:cacheprovider
run memcheck_lex 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_linear.py::test_vcycle_bit_exact" -m gpu -q -x --tb=short -p no:cacheprovider -k "False and 32"
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 128 --mode flip --workload tank --steps 4 --warmup 2"
ab rb128 X=0
ABARGS="--n 128 --mode flip --workload tank --steps 4 --warmup 2 --natural-ordering"
ab lex128 X=0
ABARGS="--n 256 --mode flip --workload tank --steps 3 --warmup 1 --natural-ordering"
ab lex256 X=0
ab lex256_nopdl JETCU_PDL=0
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s n %4d total %8.2f press %8.2f solve %8.2f iters %s" % (d["tag"], d["n"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"]))
PY
tail -3 $OUT/ab.err
