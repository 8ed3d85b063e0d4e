#!/bin/bash
# Round 2, GPU call 7 (one B200): first hardware run of the ICCG tests (full -m gpu suite), bench of HEAD, ncu of the move kernel.
set -u
OUT=gpurun_out/r2c7
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8
run bench 600 python bench.py --steps 10 --warmup 3
run ncu_move 300 ncu --set full --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum --clock-control none --import-source on -k regex:k_flip_move -c 1 -o $OUT/move python scripts/ncu_step.py 256 1
