#!/usr/bin/env python
"""Average device time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/ncu_launches.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki].split("(")[0][-48:]].append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = v[len(v) // 2:] if len(v) > 4 else v  # skip the cold first half
    print(f"{k:50s} n={len(v):4d}  mean(us)={sum(v2) / len(v2) / 1e3:9.2f}  total(us)={sum(v) / 1e3:10.1f}")
