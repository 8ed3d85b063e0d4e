#!/bin/bash
# Round 2, GPU call 15 (one B200): programmatic dependent launches for the SOR colour passes (JETCU_PDL=1), smaller wavefront
# chunks that fit the L2 (JETCU_RELAX_CHUNK / _SKEW) with and without them, particle-sector prefetch in the staggered P2G.
set -u
OUT=gpurun_out/r2c15
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
JETCU_PDL=1 run pytest_linear_pdl 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py -m gpu -q -x --tb=short -p no:cacheprovider
JETCU_PDL=1 JETCU_RELAX_CHUNK=8 run pytest_linear_pdl_chunk8 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
ab base JETCU_PDL=0
ab pdl JETCU_PDL=1
ab nopf JETCU_P2G_PREFETCH=0
ab c32s20 JETCU_RELAX_CHUNK=32 JETCU_RELAX_SKEW=20
ab c32s20_pdl JETCU_RELAX_CHUNK=32 JETCU_RELAX_SKEW=20 JETCU_PDL=1
ab c24s20_pdl JETCU_RELAX_CHUNK=24 JETCU_RELAX_SKEW=20 JETCU_PDL=1
ab c16s20_pdl JETCU_RELAX_CHUNK=16 JETCU_RELAX_SKEW=20 JETCU_PDL=1
ab c40s20_pdl JETCU_RELAX_CHUNK=40 JETCU_RELAX_SKEW=20 JETCU_PDL=1
ab c32s10_pdl JETCU_RELAX_CHUNK=32 JETCU_RELAX_SKEW=10 JETCU_PDL=1
ab c24s20 JETCU_RELAX_CHUNK=24 JETCU_RELAX_SKEW=20
ABARGS="--n 256 --mode apic --workload tank --steps 4 --warmup 2"
ab apic_base JETCU_PDL=0
ab apic_nopf JETCU_P2G_PREFETCH=0
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s %-5s %-5s total %7.2f bin %5.2f p2g %6.2f sdf %5.2f press %6.2f solve %6.2f move %6.2f iters %s" % (d["tag"], d["mode"], d["workload"], d["ms_total"], d["ms_bin"], d["ms_p2g"], d["ms_sdf"], d["ms_pressure"], d["ms_mgpcg_solve"], d["ms_move"], d["iters"]))
PY
tail -3 $OUT/ab.err
