#!/usr/bin/env python
"""Hot source lines of one kernel: scripts/ncu_hot.py prof.ncu-rep <kernel-regex> [top_n]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, lines = "", {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] not in ("", "Line No"):
        try:
            key, inst, samp = (cur, int(r[0])), int(r[7]), int(r[6])
        except ValueError:
            continue
        if key not in lines:
            lines[key] = (inst, samp, r[1].strip())
tot = sum(v[0] for v in lines.values()) or 1
ts = sum(v[1] for v in lines.values()) or 1
print("total warp inst", tot, "samples", ts)
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[1] / ts:5.1f}% samp {k[0]}:{k[1]}  {v[2][:100]}")
