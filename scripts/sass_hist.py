#!/usr/bin/env python
"""SASS opcode histogram + stall-sample share from an ncu report: sass_hist.py <file.ncu-rep> [top]"""
import csv, subprocess, sys
from collections import Counter
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, ie, ss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
c, st = Counter(), Counter()
for r in rows[2:]:
    try: n = int(r[ie]); s = int(r[ss])
    except Exception: continue
    t = r[si].split()
    if not t: continue
    op = t[1] if t[0].startswith('@') else t[0]
    key = '.'.join(op.split('.')[:3])
    c[key] += n; st[key] += s
tot, tots = sum(c.values()), max(sum(st.values()), 1)
print('total warp inst', tot)
for k, v in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print('%-28s %10d %5.1f%%  stall-samples %5.1f%%' % (k, v, 100 * v / tot, 100 * st[k] / tots))
