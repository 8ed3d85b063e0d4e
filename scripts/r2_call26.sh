#!/bin/bash
# Round 2, GPU call 26 (one B200): occupancy x fast-path variants of the SOR colour pass (JETCU_RELAX_VARIANT).
set -u
OUT=gpurun_out/r2c26
mkdir -p $OUT
ab() { local tag=$1; shift; env "$@" timeout 200 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 5 --warmup 2"
ab base JETCU_RELAX_VARIANT=0
ab v1_occ8 JETCU_RELAX_VARIANT=1
ab v2_occ8_fast JETCU_RELAX_VARIANT=2
ab v3_occ6_fast JETCU_RELAX_VARIANT=3
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s n %4d total %8.2f press %8.2f solve %8.2f iters %s" % (d["tag"], d["n"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"]))
PY
tail -3 $OUT/ab.err
