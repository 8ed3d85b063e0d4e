#!/bin/bash
# Round 2, GPU call 4 (one B200): staged P2G, chunk-interleaved SOR sweeps A/B, ncu of both.
set -u
OUT=gpurun_out/r2c4
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -3 $OUT/$name.log | tr '\n' ' ' | cut -c1-500)"; }
run pytest_gpu 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run bench 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_apic 600 python bench.py --steps 5 --warmup 3 --mode apic --no-cpu-baseline --no-projection
for c in 16 32 64; do
  JETCU_RELAX_CHUNK=$c run bench_chunk$c 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
done
JETCU_RELAX_CHUNK=32 run pytest_chunk32 900 python -m pytest tests/test_gpu_pressure.py tests/test_gpu_linear.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider
K='regex:k_p2g_tile|k_relax_colour_code'
run ncu_a 600 ncu --set full --clock-control none --import-source on -k "$K" -c 14 -f -o $OUT/p2g_relax python scripts/ncu_step.py 256 1
JETCU_RELAX_CHUNK=32 run ncu_b 600 ncu --set full --clock-control none -k regex:k_relax_colour_code -c 24 -f -o $OUT/relax_chunk32 python scripts/ncu_step.py 256 1
python scripts/ncu_summary.py rep $OUT/p2g_relax.ncu-rep > $OUT/p2g_relax_summary.txt 2>&1
python scripts/ncu_summary.py rep $OUT/relax_chunk32.ncu-rep > $OUT/relax_chunk32_summary.txt 2>&1
for f in bench bench_apic bench_chunk16 bench_chunk32 bench_chunk64; do python - $OUT/$f.log <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms'].items() if isinstance(v,float)}, d.get('fixed_work',{}).get('ms_per_pcg_iteration'), (d.get('projection') or {}).get('ms_per_iteration'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
ls -la $OUT
