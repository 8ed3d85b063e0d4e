#!/bin/bash
# Round 2, GPU call 13 (one B200): staggered P2G with pipelined loads / branch-free fractions — register-bound variants.
set -u
OUT=gpurun_out/r2c13
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_particles 600 python -m pytest tests/test_gpu_particles.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
for m in 2 3 4; do ab flip_minb$m JETCU_P2G_MINB=$m; done
ABARGS="--n 256 --mode apic --workload tank --steps 4 --warmup 2"
for m in 2 3; do ab apic_minb$m JETCU_P2G_MINB=$m; done
ABARGS="--n 256 --mode flip --workload slosh --steps 4 --warmup 2"
for m in 3 4; do ab slosh_minb$m JETCU_P2G_MINB=$m; done
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s %-5s %-5s total %7.2f bin %5.2f p2g %6.2f sdf %5.2f press %6.2f move %6.2f iters %s" % (d["tag"], d["mode"], d["workload"], d["ms_total"], d["ms_bin"], d["ms_p2g"], d["ms_sdf"], d["ms_pressure"], d["ms_move"], d["iters"]))
PY
tail -3 $OUT/ab.err
run ncu_p2g 300 ncu --set full --clock-control none --import-source on -k regex:k_p2g_stag -c 1 -o $OUT/p2g python scripts/ncu_step.py 256 1
