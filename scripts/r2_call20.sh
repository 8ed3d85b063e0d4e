#!/bin/bash
# Round 2, GPU call 20 (one B200): the whole -m gpu suite, smoke(), and the contract bench on the final build.
set -u
OUT=gpurun_out/r2c20
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; local t0=$SECONDS; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $((SECONDS-t0))s $(tail -3 $OUT/$name.log | tr '\n' ' ' | cut -c1-600)"; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider --durations=8
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_flip 600 python bench.py
run bench_apic 400 python bench.py --mode apic --steps 6 --warmup 3 --no-cpu-baseline --no-projection
run bench_pic 400 python bench.py --mode pic --steps 6 --warmup 3 --no-cpu-baseline --no-projection
