#!/bin/bash
# Round 2, GPU call 30 (one B200): two cells per thread in the SOR colour pass (JETCU_RELAX_ILP=2) — parity, digests, timing.
set -u
OUT=gpurun_out/r2c30
mkdir -p $OUT
JETCU_RELAX_ILP=2 timeout 300 python -m pytest tests/test_gpu_linear.py tests/test_gpu_pressure.py -m gpu -q -x --tb=short -p no:cacheprovider > $OUT/pytest_linear.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest_linear.log)"
ab() { local tag=$1; shift; env "$@" timeout 200 python scripts/ab_step.py --tag "$tag" $ABARGS >> $OUT/ab.jsonl 2>> $OUT/ab.err; }
ABARGS="--n 256 --mode flip --workload tank --steps 3 --warmup 2"
ab ilp2 AB_HASH=1 JETCU_RELAX_ILP=2
ab ilp1 AB_HASH=1 JETCU_RELAX_ILP=1
python - $OUT/ab.jsonl <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1])]
for d in rows:
    print("%-12s total %8.2f press %8.2f solve %8.2f iters %s %s %s %s" % (d["tag"], d["ms_total"], d["ms_pressure"], d["ms_mgpcg_solve"], d["iters"], d.get("sha_pos", "")[:10], d.get("sha_vel", "")[:10], d.get("sha_pressure", "")[:10]))
print("bit-identical:", all(rows[0][k] == rows[1][k] for k in ("sha_pos", "sha_vel", "sha_pressure")))
PY
tail -3 $OUT/ab.err
