#!/bin/bash
# Round 2, GPU call 3 (one B200): full suite with the 256^3 goldens, bench (FLIP + APIC), reference arm on the box's cores, ncu.
set -u
OUT=gpurun_out/r2c3
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -3 $OUT/$name.log | tr '\n' ' ' | cut -c1-700)"; }
nproc; free -g | head -2
run pytest_gpu 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --durations=8
run ref_arm 900 python bench.py --impl reference --steps 2 --warmup 0
run bench 900 python bench.py --steps 10 --warmup 3
run bench_apic 600 python bench.py --steps 5 --warmup 3 --mode apic --no-cpu-baseline --no-projection
run bench_pic 600 python bench.py --steps 5 --warmup 3 --mode pic --no-cpu-baseline --no-projection
run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python scripts/ncu_step.py 256 2
python scripts/ncu_summary.py launches $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
K='regex:k_p2g_tile|k_flip_move|k_pack_pairs|k_gather_attrs|k_scatter_perm|k_sort_cells|k_build_sdf'
run ncu_particles 900 ncu --set full --clock-control none --import-source on -k "$K" -c 16 -f -o $OUT/particles python scripts/ncu_step.py 256 1
python scripts/ncu_summary.py rep $OUT/particles.ncu-rep > $OUT/particles_summary.txt 2>&1
ls -la $OUT
