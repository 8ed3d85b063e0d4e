#!/bin/bash
# Round 2, GPU call 1 (one B200): reproduce the round-1 red tests un-masked, baseline bench, ncu of the particle kernels.
#   gpurun --timeout 1700 -- 'bash scripts/r2_call1.sh'
set -u
OUT=gpurun_out/r2c1
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -3 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)"; }

run t_collider 600 python -m pytest tests/test_gpu_collider.py -m gpu -q -x --tb=short -p no:cacheprovider
run t_collider_all 600 python -m pytest tests/test_gpu_collider.py -m gpu -q --tb=line -p no:cacheprovider
run t_dropin_next 600 python -m pytest tests/test_gpu_dropin_next.py -m gpu -q --tb=short -p no:cacheprovider
for e in 1 3 6; do
  mkdir -p /tmp/app$e
  run app_e$e 300 adapters/_build/hybrid_liquid_sim_cuda -e $e -r 32 -f 3 -o /tmp/app$e -m pos --compare
done
run t_smoke 600 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_fuzz.py -m gpu -q --tb=short -p no:cacheprovider
run t_rest 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_collider.py --deselect tests/test_gpu_dropin_next.py --deselect tests/test_gpu_smoke.py --deselect tests/test_gpu_fuzz.py
run bench 400 python bench.py --steps 5 --warmup 3
cp $OUT/bench.log $OUT/bench.json 2>/dev/null
run ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python scripts/ncu_step.py 256 2
python scripts/ncu_summary.py launches $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
K='regex:k_p2g_cells|k_p2g_normalise|k_move|k_gather_attrs|k_build_sdf|k_g2p|k_cell_count|k_scatter_perm|k_sort_cells'
run ncu_particles 600 ncu --set full --metrics smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --import-source on -k "$K" -c 12 -f -o $OUT/particles python scripts/ncu_step.py 256 1
python scripts/ncu_summary.py rep $OUT/particles.ncu-rep > $OUT/particles_summary.txt 2>&1
ls -la $OUT
