#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. usage: ncu_launches.py launches.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki].split("(")[0].split("<")[0].split()[-1]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':24s} {'launches':>8s} {'sum_us':>10s} {'mean_us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:24s} {len(v):8d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.1f} {sum(v) / tot * 100:6.1f}%")
