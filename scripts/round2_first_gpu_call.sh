#!/bin/bash
# First GPU call of round 2 (one B200): everything round 1 could not run after its GPU minutes were gone.
#   gpurun --timeout 1500 -- 'bash scripts/round2_first_gpu_call.sh'
# Every step has its own timeout; results land in gpurun_out/r2_first/.
set -u
OUT=gpurun_out/r2_first
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-300)"; }

# 1. the established parity suite + the first-run files in child processes (tests/test_gpu_zz_firstrun.py: XPASS = the new
#    CUDA paths work; their logs land in gpurun_out/firstrun/)
run pytest_gpu 1500 python -m pytest tests -m gpu -q -rxX
# 2. contract bench (graphs on) and the smoke probe
run bench 400 python bench.py --steps 5 --warmup 3
run bench_smoke 300 python scripts/bench_smoke.py 256 5 3
# 3. P2G variants A/B: stage timers of the same 5 steps (ms_p2g in stages_ms)
for k in default carryfree serial particle; do
  JETCU_P2G_KERNEL=$k run bench_p2g_$k 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-projection
done
# 3b. L2-resident chunk-interleaved SOR sweep A/B (bit-identical by construction and by tests/test_hostcheck_linear.py):
#     mgpcg_solve_ms / stages_ms.ms_pressure and projection.mgpcg_solve_ms of the same bench
for c in 16 32 64; do
  JETCU_RELAX_CHUNK=$c run bench_relax_chunk_$c 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
done
JETCU_RELAX_CHUNK=32 run pytest_gpu_chunk32 400 python -m pytest tests/test_gpu_pressure.py tests/test_gpu_linear.py -m gpu -q
# 4. ncu: launch list of one sub-step and full captures of the kernels the step spends its time in
run ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python scripts/ncu_step.py 256 2
for k in k_relax_colour_code k_p2g_cells k_move k_gather_attrs k_mvm; do
  run ncu_$k 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -f -o $OUT/$k python scripts/ncu_step.py 256 1
done
python scripts/ncu_summary.py launches $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
for k in k_relax_colour_code k_p2g_cells k_move k_gather_attrs k_mvm; do
  python scripts/ncu_summary.py rep $OUT/$k.ncu-rep > $OUT/${k}_summary.txt 2>&1
done
ls -la $OUT

# 5. multi-GPU (separate calls, --gpus 2 then 8): scripts/round2_multi_gpu_call.sh — slab parity with both halo transports, halo
#    latency, the slab-plan and peer-memory A/B benches.
