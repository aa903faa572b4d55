#!/usr/bin/env python3
"""Per-source-line warp instruction counts of one kernel from an ncu report captured with --import-source on.
usage: ncu_src_lines.py rep.ncu-rep launch_index [top_n] [source_file_for_text]"""
import csv, io, subprocess, sys, collections
rep, idx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
srcfile = sys.argv[4] if len(sys.argv) > 4 else "opendlv-perception-vision-orbslam2_b200/csrc/orbx_kernels.cu"
try:
    text = open(srcfile).read().splitlines()
except OSError:
    text = []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
per = collections.defaultdict(lambda: [0, 0, 0])
seen = set()
line = None
tot = 0
ops = collections.Counter()
for r in rows[hdr + 1:]:
    if not r: continue
    if r[0] != "":
        line = r[0]; continue
    if len(r) < 8 or not r[2].startswith("0x"): continue
    if r[2] in seen: continue
    seen.add(r[2])
    try: n = int(float(r[7])); s = int(float(r[6]))
    except ValueError: continue
    per[line][0] += n; per[line][1] += 1; per[line][2] += s
    tot += n
    op = r[3].split()
    op = [t for t in op if not t.startswith("@")][0].split(".")[0]
    ops[op] += n
print(f"total warp instructions {tot}")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    t = text[int(k) - 1].strip()[:120] if k and k.isdigit() and int(k) <= len(text) else ""
    print(f"{v[0]:>10d} {100 * v[0] / tot:5.1f}%  sass={v[1]:4d} samples={v[2]:6d}  L{k}: {t}")
print("opcode mix:", ", ".join(f"{o}={100 * n / tot:.1f}%" for o, n in ops.most_common(25)))
