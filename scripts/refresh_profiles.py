#!/usr/bin/env python3
"""Turn the raw outputs of a profiling gpurun (gpurun_out/launches_X.csv, prof_all_X.ncu-rep, bench_X.json) into the
tracked summaries under profiles/ (named rN_*, N = round).  usage: refresh_profiles.py X [round]   (e.g. r2 2)"""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = "r" + (sys.argv[2] if len(sys.argv) > 2 else "1")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True, check=True).stdout
open(os.path.join(P, f"{rnd}_launch_summary.txt"), "w").write(run(os.path.join(ROOT, "scripts", "ncu_launches.py"), os.path.join(G, f"launches_{tag}.csv")))
full = run(os.path.join(ROOT, "scripts", "ncu_raw.py"), os.path.join(G, f"prof_all_{tag}.ncu-rep"))
open(os.path.join(P, f"{rnd}_ncu_full_summary.txt"), "w").write(full)
shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{rnd}_launches.csv"))
shutil.copy(os.path.join(G, f"bench_{tag}.json"), os.path.join(P, f"{rnd}_final_bench.json"))
names = {"k_copy_level0": "copy_level0", "k_fast_segs": "fast", "k_resize": "resize", "k_blur": "blur", "k_octree": "octree", "k_describe": "describe"}
byt, win, cnt = {}, {}, {}
for line in full.strip().split("\n"):
    d = dict(kv.split("=") for kv in line.split())
    k = names[d["kernel"]]
    byt[k] = byt.get(k, 0) + float(d["rdMB"]) + float(d["wrMB"]); win[k] = win.get(k, 0) + float(d["winst"]); cnt[k] = cnt.get(k, 0) + 1
json.dump({"source": f"profiles/{rnd}_ncu_full_summary.txt (ncu --set full --clock-control none, one 64-frame step, ORBX_SPLIT=1 so each stage is one "
                     "launch per level group; dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum summed over the stage's launches)",
           "bytes_per_step": {k: int(v * 1e6) for k, v in byt.items()}, "launches_per_step": cnt,
           "warp_instructions_per_step": win}, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{rnd}_launch_summary.txt")).read())
print({k: round(v, 1) for k, v in byt.items()}, sum(byt.values()), {k: f"{v:.3g}" for k, v in win.items()}, f"{sum(win.values()):.4g}")
