#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
tot = 0.0
n = 0
for idx, row in enumerate(r):
    if idx < skip:
        continue
    v = float(row[vi].replace(',', ''))
    u = row[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    short = re.sub(r'\(.*', '', re.sub(r'<.*', '', row[ki])).replace('msm::', '').replace('void ', '')
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    n += 1
print(f"total {tot:.1f} us over {n} launches")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:22s} n={c:4d} total={v:10.1f} us avg={v / c:9.1f} us share={v / tot * 100:5.1f}%")
