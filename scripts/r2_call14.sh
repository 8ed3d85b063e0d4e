#!/bin/bash
# Round 2, GPU call 14 (one B200): row-major particle bins (JETCU_BIN_ORDER=rows: the particles of one octant of an x-row of
# cells are contiguous -> coalesced particle loads in the staggered P2G) against cell-major bins.
set -u
OUT=gpurun_out/r2c14
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
JETCU_BIN_ORDER=rows run pytest_particles_rows 600 python -m pytest tests/test_gpu_particles.py tests/test_gpu_collider.py -m gpu -q -x --tb=short -p no:cacheprovider
ab() { local tag=$1; shift; env "$@" timeout 300 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 6 --warmup 3"
ab flip_cells JETCU_BIN_ORDER=cells
ab flip_rows JETCU_BIN_ORDER=rows
ABARGS="--n 256 --mode apic --workload tank --steps 4 --warmup 2"
ab apic_cells JETCU_BIN_ORDER=cells
ab apic_rows JETCU_BIN_ORDER=rows
ABARGS="--n 256 --mode flip --workload slosh --steps 4 --warmup 2"
ab slosh_cells JETCU_BIN_ORDER=cells
ab slosh_rows JETCU_BIN_ORDER=rows
ABARGS="--n 256 --mode flip --workload dam --steps 4 --warmup 2"
ab dam_cells JETCU_BIN_ORDER=cells
ab dam_rows JETCU_BIN_ORDER=rows
ABARGS="--n 256 --mode pic --workload slosh --steps 4 --warmup 2"
ab picslosh_cells JETCU_BIN_ORDER=cells
ab picslosh_rows JETCU_BIN_ORDER=rows
python - $OUT/ab.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print("%-14s %-5s %-5s total %7.2f bin %5.2f p2g %6.2f sdf %5.2f press %6.2f move %6.2f iters %s" % (d["tag"], d["mode"], d["workload"], d["ms_total"], d["ms_bin"], d["ms_p2g"], d["ms_sdf"], d["ms_pressure"], d["ms_move"], d["iters"]))
PY
tail -3 $OUT/ab.err
JETCU_BIN_ORDER=rows run ncu_p2g 300 ncu --set full --clock-control none --import-source on -k regex:k_p2g_stag -c 1 -o $OUT/p2g_rows python scripts/ncu_step.py 256 1
