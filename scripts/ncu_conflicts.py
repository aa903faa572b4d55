#!/usr/bin/env python3
"""Shared-memory wavefronts (actual / ideal) and stall samples per SASS instruction of one launch.
usage: ncu_conflicts.py report.ncu-rep kernel_regex launch_index [min_excess]"""
import csv, subprocess, sys, io
rep, kern, idx = sys.argv[1], sys.argv[2], int(sys.argv[3])
thr = int(sys.argv[4]) if len(sys.argv) > 4 else 20000
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][-1]
hdr = rows[h]
ix = {c: hdr.index(c) for c in ("Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive", "# Samples")}
body = [r for r in rows[h + 1:] if r and r[0].startswith("0x")]
tw = ti = 0
for j, r in enumerate(body):
    w, i, e = (int(r[ix[c]] or 0) for c in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive"))
    tw += w; ti += i
    if e >= thr:
        print(f"{j:5d} inst {int(r[ix['Instructions Executed']]):8d} wave {w:8d} ideal {i:8d} x{w / max(i, 1):.2f} smp {r[ix['# Samples']]:>4}  {r[1].strip()[:60]}")
print("total wavefronts", tw, "ideal", ti)
