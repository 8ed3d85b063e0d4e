#!/bin/bash
# Round 2, GPU call 5 (one B200): quadrant P2G (octant-grouped cells), skewed SOR wavefront A/B, ncu (reports stay on the box).
set -u
OUT=gpurun_out/r2c5
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -3 $OUT/$name.log | tr '\n' ' ' | cut -c1-500)"; }
run pytest_gpu 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
run bench 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_apic 600 python bench.py --steps 5 --warmup 3 --mode apic --no-cpu-baseline --no-projection
for cs in "32 5" "16 5" "64 5" "32 10" "24 4"; do set -- $cs
  JETCU_RELAX_CHUNK=$1 JETCU_RELAX_SKEW=$2 run bench_skew_c$1_s$2 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
done
JETCU_RELAX_CHUNK=32 JETCU_RELAX_SKEW=5 run pytest_skew 900 python -m pytest tests/test_gpu_pressure.py tests/test_gpu_linear.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider
K='regex:k_p2g_tile|k_sort_cells|k_scatter_perm|k_gather_attrs|k_cell_count'
run ncu_a 600 ncu --set full --clock-control none --import-source on -k "$K" -c 14 -f -o /tmp/p2g python scripts/ncu_step.py 256 1
python scripts/ncu_summary.py rep /tmp/p2g.ncu-rep > $OUT/p2g_summary.txt 2>&1
ncu -i /tmp/p2g.ncu-rep --page raw --csv > /tmp/p2g_raw.csv 2>/dev/null
python - > $OUT/p2g_detail.txt 2>&1 <<'PY'
import csv
rows=list(csv.reader(open('/tmp/p2g_raw.csv')))
hdr,units=rows[0],rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
seen=set()
for r in rows[2:]:
    name=r[hdr.index('Kernel Name')][:28]
    if name in seen: continue
    seen.add(name)
    print('=====',name)
    for w in keys:
        if w in hdr: print('  %-80s %s %s'%(w,r[hdr.index(w)],units[hdr.index(w)]))
    st=[]
    for i,w in enumerate(hdr):
        if w.startswith('smsp__average_warps_issue_stalled') and w.endswith('per_issue_active.ratio'):
            try: st.append((float(r[i]), w.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except: pass
    st.sort(reverse=True)
    print('  stalls:', ', '.join('%s %.2f'%(b,a) for a,b in st[:7]))
PY
for f in $OUT/bench*.log; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms'].items() if isinstance(v,float)}, d.get('fixed_work',{}).get('ms_per_pcg_iteration'), (d.get('projection') or {}).get('ms_per_iteration'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
ls -la $OUT; du -sh gpurun_out
