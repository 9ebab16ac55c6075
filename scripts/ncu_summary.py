#!/usr/bin/env python
"""Summarise an .ncu-rep (no GPU needed): scripts/ncu_summary.py prof.ncu-rep [kernel-substring]
Per kernel: duration, DRAM bytes, issue/fp64 utilisation, occupancy, instruction mix, stall reasons."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio",
       "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for r in data:
    name = r[col["Kernel Name"]]
    if filt not in name:
        continue
    print("=" * 100)
    print(name[:140])
    for k in KEY:
        if k in col:
            print(f"  {k:70s} {r[col[k]]:>16s} {units[col[k]]}")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("  stalls (warps per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))
    pipes = []
    for h, i in col.items():
        if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum"):
            try:
                pipes.append((float(r[i]), h[len("sm__inst_executed_pipe_"):-4]))
            except ValueError:
                pass
    pipes.sort(reverse=True)
    print("  pipes (warp inst): " + ", ".join(f"{n}={v:.3g}" for v, n in pipes[:10]))
