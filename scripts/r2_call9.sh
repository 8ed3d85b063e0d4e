#!/bin/bash
# Round 2, GPU call 9 (one B200): full -m gpu suite on the build with the new P2G / advection defaults and the diffusion solver;
# bench of HEAD (FLIP, then APIC and PIC quick lines); ncu of the P2G tile kernel as it now runs.
set -u
OUT=gpurun_out/r2c9
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=5
run bench 600 python bench.py --steps 10 --warmup 3
run bench_apic 300 python bench.py --steps 6 --warmup 3 --mode apic --no-cpu-baseline --no-projection --e2e-steps 0
run bench_pic 300 python bench.py --steps 6 --warmup 3 --mode pic --no-cpu-baseline --no-projection --e2e-steps 0
run ncu_p2g 300 ncu --set full --clock-control none --import-source on -k regex:k_p2g_tile -c 2 -o $OUT/p2g python scripts/ncu_step.py 256 1
