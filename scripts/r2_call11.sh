#!/bin/bash
# Round 2, GPU call 11 (one B200): conversion-free barycentric (no FRND / F2I / I2F on the XU pipe) — A/B of the advection
# kernel variants on top of it, all three modes.
set -u
OUT=gpurun_out/r2c11
mkdir -p $OUT
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
for c in 0 2 3 4 5; do ab move_cc$c JETCU_MOVE_CACHE=$c; done
ABARGS="--n 256 --mode pic --workload tank --steps 4 --warmup 2"
for c in 0 2 4 5; do ab pic_cc$c JETCU_MOVE_CACHE=$c; done
ABARGS="--n 256 --mode apic --workload tank --steps 4 --warmup 2"
for c in 0 2 4; do ab apic_cc$c JETCU_MOVE_CACHE=$c; done
ABARGS="--n 256 --mode flip --workload slosh --steps 4 --warmup 2"
for c in 0 2 5; do ab slosh_cc$c JETCU_MOVE_CACHE=$c; done
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s %-5s %-5s total %7.2f bin %5.2f p2g %6.2f sdf %5.2f press %6.2f move %6.2f iters %s" % (d["tag"], d["mode"], d["workload"], d["ms_total"], d["ms_bin"], d["ms_p2g"], d["ms_sdf"], d["ms_pressure"], d["ms_move"], d["iters"]))
PY
tail -3 $OUT/ab.err
