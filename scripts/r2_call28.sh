#!/bin/bash
# Round 2, GPU call 28 (one B200): colour-split storage of the finest level's SOR sweeps — bit identity against the natural
# layout (digests of positions, velocities and pressure after 5 sub-steps at 256^3) and timing.
set -u
OUT=gpurun_out/r2c28
mkdir -p $OUT
ab() { local tag=$1; shift; env "$@" timeout 200 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 3 --warmup 2"
ab split AB_HASH=1 JETCU_RELAX_SPLIT=1
ab natural AB_HASH=1 JETCU_RELAX_SPLIT=0
ABARGS="--n 256 --mode flip --workload dam --steps 2 --warmup 2"
ab split_dam AB_HASH=1 JETCU_RELAX_SPLIT=1
ab natural_dam AB_HASH=1 JETCU_RELAX_SPLIT=0
python - $OUT/ab.jsonl <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1])]
for d in rows:
    print("%-12s total %8.2f press %8.2f solve %8.2f iters %s %s %s %s" % (d["tag"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"], d.get("sha_pos", "")[:10], d.get("sha_vel", "")[:10], d.get("sha_pressure", "")[:10]))
for a, b in ((0, 1), (2, 3)):
    if len(rows) > b:
        print(rows[a]["tag"], "vs", rows[b]["tag"], "bit-identical:", all(rows[a][k] == rows[b][k] for k in ("sha_pos", "sha_vel", "sha_pressure")))
PY
tail -3 $OUT/ab.err
