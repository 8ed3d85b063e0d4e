#!/bin/bash
# Round 2, GPU call 2 (one B200): full -m gpu suite (no xfail), bench with the atomic-free P2G, scene exploration, ncu of k_p2g_tile.
set -u
OUT=gpurun_out/r2c2
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -3 $OUT/$name.log | tr '\n' ' ' | cut -c1-600)"; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run bench 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run explore 900 python scripts/explore_scenes.py 256
K='regex:k_p2g_tile|k_p2g_normalise|k_gather_attrs'
run ncu_p2g 600 ncu --set full --clock-control none --import-source on -k "$K" -c 12 -f -o $OUT/p2g python scripts/ncu_step.py 256 1
python scripts/ncu_summary.py rep $OUT/p2g.ncu-rep > $OUT/p2g_summary.txt 2>&1
ls -la $OUT
