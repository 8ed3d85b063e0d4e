#!/bin/bash
# Round 2, GPU call 12 (one B200): staggered-lattice P2G (k_p2g_stag) on (cell, octant)-sorted particles — full -m gpu suite,
# A/B against the 18-slot window kernel in all three modes, ncu of the new kernel.
set -u
OUT=gpurun_out/r2c12
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_particles 600 python -m pytest tests/test_gpu_particles.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
ab flip_stag JETCU_P2G_KERNEL=stag
ab flip_window JETCU_P2G_KERNEL=window
ABARGS="--n 256 --mode apic --workload tank --steps 4 --warmup 2"
ab apic_stag JETCU_P2G_KERNEL=stag
ab apic_window JETCU_P2G_KERNEL=window
ABARGS="--n 256 --mode flip --workload slosh --steps 4 --warmup 2"
ab slosh_stag JETCU_P2G_KERNEL=stag
ab slosh_window JETCU_P2G_KERNEL=window
ABARGS="--n 256 --mode flip --workload dam --steps 4 --warmup 2"
ab dam_stag JETCU_P2G_KERNEL=stag
ab dam_window JETCU_P2G_KERNEL=window
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s %-5s %-5s total %7.2f bin %5.2f p2g %6.2f sdf %5.2f press %6.2f move %6.2f iters %s" % (d["tag"], d["mode"], d["workload"], d["ms_total"], d["ms_bin"], d["ms_p2g"], d["ms_sdf"], d["ms_pressure"], d["ms_move"], d["iters"]))
PY
tail -3 $OUT/ab.err
run pytest_gpu 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=5 --deselect tests/test_gpu_particles.py
run ncu_p2g 300 ncu --set full --clock-control none --import-source on -k regex:k_p2g_stag -c 2 -o $OUT/p2g python scripts/ncu_step.py 256 1
