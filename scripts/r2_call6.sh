#!/bin/bash
# Round 2, GPU call 6 (one B200): SOR wavefront sweep of (chunk, skew), P2G tile height A/B, ncu of the relax kernel (traffic).
set -u
OUT=gpurun_out/r2c6
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $(tail -2 $OUT/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run pytest_gpu 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x
for cs in "0 0" "64 5" "64 10" "128 5" "128 10" "96 5" "64 20" "48 5" "256 5"; do set -- $cs
  JETCU_RELAX_CHUNK=$1 JETCU_RELAX_SKEW=$2 run bench_c$1_s$2 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
done
JETCU_P2G_TZ=4 run bench_p2g_tz4 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-projection
# traffic of the relax kernel, whole passes vs wavefront (DRAM bytes per launch -> profiles/roofline_traffic.json)
JETCU_RELAX_CHUNK=0 run ncu_whole 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none -k regex:k_relax_colour_code -c 400 --csv --log-file $OUT/relax_whole.csv python scripts/ncu_step.py 256 1
run ncu_wave 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none -k regex:k_relax_colour_code -c 1400 --csv --log-file $OUT/relax_wave.csv python scripts/ncu_step.py 256 1
for f in $OUT/bench*.log; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms'].items() if isinstance(v,float)}, d.get('fixed_work',{}).get('ms_per_pcg_iteration'), (d.get('projection') or {}).get('ms_per_iteration'), d['roofline']['frac'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
python - $OUT <<'PY'
import csv, sys, os, collections
for name in ("relax_whole", "relax_wave"):
    path = os.path.join(sys.argv[1], name + ".csv")
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r and "Metric Value" in r: hdr = r; rows = rows[i+1:]; break
    kn, mn, mv, gs, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r[idc], {"grid": r[gs]})[r[mn]] = float(r[mv].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for k, d in per.items():
        a = agg[d["grid"]]; a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0); a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
    print(name)
    for g, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]:
        print("  grid %-18s launches %5d  time/launch %9.1f  dram read/launch %12.0f  write/launch %12.0f" % (g, a[0], a[1] / a[0], a[2] / a[0], a[3] / a[0]))
PY
rm -f $OUT/relax_whole.csv $OUT/relax_wave.csv
ls -la $OUT
