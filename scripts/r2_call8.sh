#!/bin/bash
# Round 2, GPU call 8 (one B200): full -m gpu suite on the build with the corner-cached advection kernel as default; A/B of the
# advection cache, the P2G loop forms / tile heights, smaller SOR wavefront chunks; APIC; slosh scene.
set -u
OUT=gpurun_out/r2c8
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=5
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
ab base JETCU_MOVE_CACHE=0
ab cache1 JETCU_MOVE_CACHE=1
ab cache2 JETCU_MOVE_CACHE=2
ab p2g_pf JETCU_P2G_PF=1
ab p2g_tz4 JETCU_P2G_TZ=4
ab p2g_tz4_pf JETCU_P2G_TZ=4 JETCU_P2G_PF=1
ab sor_32_20 JETCU_RELAX_CHUNK=32 JETCU_RELAX_SKEW=20
ab sor_24_20 JETCU_RELAX_CHUNK=24 JETCU_RELAX_SKEW=20
ab sor_40_20 JETCU_RELAX_CHUNK=40 JETCU_RELAX_SKEW=20
ABARGS="--n 256 --mode flip --workload slosh --steps 4 --warmup 2"
ab slosh_base JETCU_MOVE_CACHE=0
ab slosh_cache1 JETCU_MOVE_CACHE=1
ABARGS="--n 256 --mode apic --workload tank --steps 4 --warmup 2"
ab apic_base JETCU_MOVE_CACHE=1
ab apic_pf JETCU_P2G_PF=1
ab apic_minb2 JETCU_P2G_MINB=2
ABARGS="--n 256 --mode pic --workload tank --steps 4 --warmup 2"
ab pic_base JETCU_MOVE_CACHE=0
ab pic_cache1 JETCU_MOVE_CACHE=1
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s %-5s %-5s total %7.2f bin %5.2f p2g %6.2f sdf %5.2f press %6.2f move %6.2f iters %s" % (d["tag"], d["mode"], d["workload"], d["ms_total"], d["ms_bin"], d["ms_p2g"], d["ms_sdf"], d["ms_pressure"], d["ms_move"], d["iters"]))
PY
tail -3 $OUT/ab.err
