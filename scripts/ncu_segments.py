#!/usr/bin/env python3
"""Per-barrier-segment instruction shares of one kernel launch in an ncu report (SASS page).
usage: ncu_segments.py report.ncu-rep kernel_regex launch_index [dump.txt]"""
import csv, subprocess, sys, io
rep, kern, idx = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][-1]
hdr = rows[h]; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
body = [r for r in rows[h + 1:] if r and r[0].startswith("0x")]
tot = sum(int(r[iI]) for r in body)
print("total warp instructions", tot, "sass lines", len(body))
seg = smp = 0; first = 0
for j, r in enumerate(body):
    seg += int(r[iI]); smp += int(r[iS])
    if "BAR.SYNC" in r[1] or j == len(body) - 1:
        print(f"sass {first:5d}-{j:5d}: {100 * seg / tot:5.1f}% inst ({seg:10d}), samples {smp}")
        seg = smp = 0; first = j + 1
if len(sys.argv) > 4:
    open(sys.argv[4], "w").write("\n".join(f"{j:5d} {int(r[iI]):9d} {int(r[iS]):5d} {r[1]}" for j, r in enumerate(body)))
