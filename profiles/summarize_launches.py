#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]
ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki])
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.0f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} n={v[0]:4d} total={v[1]:10.1f} us avg={v[1] / v[0]:8.1f} share={100 * v[1] / tot:5.1f}%")
