#!/usr/bin/env python3
"""Static SASS mnemonic census per kernel of the sm_100a cubins in csrc/_build (evidence for the design claims:
TMA / mbarrier / cp.async / packed min-max / dp4a / ballot-match-shuffle / POPC, no tensor-core instructions).
usage: sass_census.py > profiles/r1_sass_census.txt      (needs cuobjdump and nvdisasm on PATH)"""
import collections, glob, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sorted(glob.glob(os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200", "csrc", "_build", "*.o")))
txt = ""
with tempfile.TemporaryDirectory() as d:
    for o in objs:
        subprocess.run(["cuobjdump", "-xelf", "all", o], cwd=d, check=True, stdout=subprocess.DEVNULL)
    for c in sorted(glob.glob(os.path.join(d, "*.cubin"))):
        txt += subprocess.run(["nvdisasm", "-c", c], capture_output=True, text=True).stdout
parts = re.split(r"//-+ \.text\.(\S+) -+", txt)
keys = ["UTMALDG", "SYNCS", "LDGSTS", "ACQBULK", "VIMNMX3", "VABSDIFF4", "IDP.4A", "IDP.2A", "PRMT", "VOTE", "SHFL", "MATCH", "REDUX",
        "POPC", "LOP3", "ATOMS", "LDS", "STS", "LDG", "STG", "HMMA", "UTCHMMA", "IMMA"]
print("SASS mnemonic census of liborbx.so (nvdisasm of the sm_100a cubins built by csrc/Makefile; static instruction counts per kernel).")
print("Evidence for the design claims: TMA box loads (UTMALDG) completing on mbarriers (SYNCS) in k_resize / k_blur / k_fast_segs /")
print("k_describe; programmatic dependent launch (ACQBULK = griddepcontrol.wait) in k_resize; packed three-input")
print("min/max (VIMNMX3.S16x2) and VABSDIFF4 in k_fast_segs; dp4a / dp2a (IDP.4A / IDP.2A) in k_blur, k_resize, k_describe; ballot /")
print("match / shuffle (VOTE, MATCH, SHFL) in k_octree; XOR + POPC on the INT pipe and no tensor-core instruction (HMMA, UTC*MMA, IMMA)")
print("in the matchers (k_knn2_*, k_distinctive, k_voc_transform, k_stereo_match).\n")
print(f"{'kernel':18s} " + " ".join(f"{k:>9s}" for k in keys) + "     total")
for i in range(1, len(parts), 2):
    name, body = parts[i], parts[i + 1]
    m = re.match(r"_Z(\d+)", name)
    short = name[2 + len(m.group(1)):2 + len(m.group(1)) + int(m.group(1))] if m else name[:18]
    ins = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
    c = collections.Counter()
    for op in ins:
        for k in keys:
            if op == k or op.startswith(k + ".") or (k not in ("LDS", "STS", "LDG", "STG") and op.startswith(k)):
                c[k] += 1
                break
    print(f"{short:18s} " + " ".join(f"{c[k]:9d}" for k in keys) + f" {len(ins):9d}")
