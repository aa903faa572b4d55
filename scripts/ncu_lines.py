#!/usr/bin/env python3
"""Per-source-line instruction share, stall-sample share and active lanes of one kernel launch in an ncu report.
usage: ncu_lines.py rep.ncu-rep kernel_regex [launch_skip] [min_pct]"""
import csv, collections, subprocess, sys, io, os
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
H = rows[hdr]
iL, iI, iS, iT = H.index("Line No"), H.index("Instructions Executed"), H.index("# Samples"), H.index("Thread Instructions Executed")
ins, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) > iI and r[iL].strip().isdigit():
        try:
            ln = int(r[iL]); ins[ln] += int(r[iI]); smp[ln] += int(r[iS]); thr[ln] += int(r[iT])
        except ValueError:
            pass
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "opendlv-perception-vision-orbslam2_b200/csrc", sys.argv[5] if len(sys.argv) > 5 else "orbx_kernels.cu")).read().split("\n")
tot, ts = sum(ins.values()) or 1, sum(smp.values()) or 1
print(f"warp-instructions {tot}  samples {ts}")
for ln in sorted(ins):
    if ins[ln] / tot * 100 >= minp or smp[ln] / ts * 100 >= minp:
        text = src[ln - 1][:100] if ln - 1 < len(src) else ""
        print(f"{ln:5d} {ins[ln] / tot * 100:5.1f}% ins {smp[ln] / ts * 100:5.1f}% smp lanes {thr[ln] / max(ins[ln], 1):4.0f} | {text}")
