#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line:
instructions executed and stall samples.  usage: ncu_src.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
H = rows[hdr]
end = next((i for i in range(hdr + 1, len(rows)) if rows[i] and rows[i][0] == "Line No"), len(rows))
iLine, iIns, iSamp = H.index("Line No"), H.index("Instructions Executed"), H.index("# Samples")
src = {}
for r in rows[:hdr]:
    if len(r) >= 2 and r[0].isdigit():
        src[int(r[0])] = r[1]
ins = collections.Counter(); samp = collections.Counter()
for r in rows[hdr + 1:end]:
    if len(r) <= iIns or not r[iLine].strip().isdigit():
        continue
    ln = int(r[iLine])
    try:
        ins[ln] += int(r[iIns]); samp[ln] += int(r[iSamp])
    except ValueError:
        pass
tot = sum(ins.values()) or 1; ts = sum(samp.values()) or 1
print(f"total warp-instructions {tot}, samples {ts}")
for ln, v in ins.most_common(top):
    print(f"{ln:5d} {v / tot * 100:5.1f}% ins {samp[ln] / ts * 100:5.1f}% smp | {src.get(ln, '')[:110]}")
