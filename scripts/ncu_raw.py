#!/usr/bin/env python3
"""Print a compact per-kernel table from `ncu -i rep --page raw --csv`. usage: ncu_raw.py rep.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "winst"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("launch__grid_size", "grid")]
idx = [(H.index(k), n) for k, n in want if k in H]
units = rows[1]
for r in rows[2:]:
    parts = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].split("<")[0].split()[-1][:16]      # "void k_blur<(bool)0>(...)": neither return type nor template arguments
        else:
            try:
                f = float(v.replace(",", ""))
                if units[i] in ("byte",): f /= 1e6
                if units[i] in ("Kbyte",): f /= 1e3
                if units[i] in ("Gbyte",): f *= 1e3
                if units[i] == "ms": f *= 1e3
                if units[i] == "ns" : f /= 1e3
                v = f"{f:.1f}" if f < 1e6 else f"{f:.3g}"
            except ValueError:
                pass
        parts.append(f"{n}={v}")
    print(" ".join(parts))
