#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo and --import-source on):
scripts/ncu_lines.py prof.ncu-rep [top_n]   -> lines sorted by warp instructions executed, with stall samples."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, lines = "", []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 10 and r[0] not in ("", "Line No"):
        try:
            lines.append((int(r[7]), int(r[6]), cur_file, int(r[0]), r[1].strip(), int(r[8])))
        except ValueError:
            pass
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[1] for l in lines) or 1
print(f"total warp inst {tot_i:.4g}, samples {tot_s}")
for inst, samp, f, ln, src, tinst in sorted(lines, reverse=True)[:top]:
    print(f"{100 * inst / tot_i:5.1f}% inst {100 * samp / tot_s:5.1f}% samp  thr/warp {tinst / max(inst, 1):4.1f}  {f}:{ln}  {src[:110]}")
