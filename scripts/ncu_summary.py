#!/usr/bin/env python
"""Summarise ncu output for profiles/:
    ncu_summary.py rep <file.ncu-rep>        -> per captured launch: duration, DRAM bytes, throughput, occupancy, regs
    ncu_summary.py launches <launches.csv>   -> per kernel: launches, total / mean duration, share of the captured time
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fp64.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
            "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]
    for r in rows[2:]:
        for w in want:
            if w in hdr:
                print("%-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        print()


def launches(path, segment=None):
    """segment = "kernel:k" keeps the launches from the k-th occurrence (1-based) of `kernel` up to the next one
    (e.g. k_cell_count:2 = the second sub-step of a bench run: every step starts with one k_cell_count)."""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = None
    for idx, r in enumerate(rows):
        if "Kernel Name" in r and "Metric Value" in r:
            hdr = r
            rows = rows[idx + 1:]
            break
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    if segment:
        name, k = segment.split(":")
        hits = [i for i, r in enumerate(rows) if r[kn].split("(")[0] == name]
        lo = hits[int(k) - 1]
        hi = hits[int(k)] if int(k) < len(hits) else len(rows)
        # a step begins two launches earlier (k_zero_walls precedes the binning) — immaterial for the shares
        rows = rows[lo:hi]
        print("# segment %s: launches %d..%d of the capture" % (segment, lo, hi))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        unit = r[mu]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
        total += us
    print("%-44s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "mean_us", "share"))
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %12.1f %10.2f %6.1f%%" % (name[:44], cnt, us, us / cnt, 100 * us / total))
    print("%-44s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), total))


if __name__ == "__main__":
    {"rep": rep, "launches": launches}[sys.argv[1]](*sys.argv[2:])
