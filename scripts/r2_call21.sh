#!/bin/bash
# Round 2, GPU call 21 (one B200): ICCG tile sweeps staged in shared memory — parity (linear, pressure, diffusion, drop-in, collider
# scenes that run ICCG) and the cost of an ICCG pressure solve at 128^3 / 256^3 before (JETCU build of call 20 = profiles) and after.
set -u
OUT=gpurun_out/r2c21
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; local t0=$SECONDS; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $((SECONDS-t0))s $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_iccg 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py tests/test_gpu_diffusion.py tests/test_gpu_dropin.py tests/test_gpu_dropin_next.py tests/test_gpu_smoke.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 128 --mode flip --workload tank --steps 3 --warmup 1 --iccg"
ab iccg128 X=0
ABARGS="--n 256 --mode flip --workload tank --steps 2 --warmup 1 --iccg"
ab iccg256 X=0
ab iccg256_nopdl JETCU_PDL=0
# the same with the previous build (tile sweeps in global memory): scripts/_old/libjetcu_prev.so = HEAD before this change
if [ -f scripts/_old/libjetcu_prev.so ]; then
  cp fluid-engine-dev_b200/libjetcu.so /tmp/libjetcu_new.so; cp scripts/_old/libjetcu_prev.so fluid-engine-dev_b200/libjetcu.so
  ABARGS="--n 128 --mode flip --workload tank --steps 3 --warmup 1 --iccg"
  ab prev_iccg128 X=0
  ABARGS="--n 256 --mode flip --workload tank --steps 2 --warmup 1 --iccg"
  ab prev_iccg256 X=0
  cp /tmp/libjetcu_new.so fluid-engine-dev_b200/libjetcu.so
fi
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s n %4d total %8.2f press %8.2f solve %8.2f iters %s" % (d["tag"], d["n"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"]))
PY
tail -3 $OUT/ab.err
