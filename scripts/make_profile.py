#!/usr/bin/env python
"""Turn one gpu_profile.sh run (gpurun_out/{bench,launches,prof}_<tag>.*) into the tracked artefacts under
profiles/: <tag>_bench.json, <tag>_launches.csv, <tag>_ncu_full_raw.csv, <tag>_ncu_summary.txt, <tag>_summary.md
and traffic.json (DRAM bytes per launch and issue-slot utilisation per kernel, read by bench.py).
usage: scripts/make_profile.py <tag> ["title line"]"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else f"profile {tag}"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rep = os.path.join(G, f"prof_{tag}.ncu-rep")


def short(name):
    m = re.match(r"(?:void )?(?:cgrb::)?(\w+)(<[^>]*>)?", name)
    base, targs = m.group(1), m.group(2) or ""
    if base in ("k_gap_prefix", "k_seg_scan"):
        return base + targs.replace("(int)", "")
    return base


shutil.copy(os.path.join(G, f"bench_{tag}.json"), os.path.join(P, f"{tag}_bench.json"))
shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv"))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_ncu_full_raw.csv"), "w").write(raw)
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True,
                      text=True).stdout
open(os.path.join(P, f"{tag}_ncu_summary.txt"), "w").write(summ)

rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, k):
    try:
        return float(r[col[k]])
    except (KeyError, ValueError):
        return float("nan")


full = {}
for r in data:
    n = short(r[col["Kernel Name"]])
    full[n] = dict(ms=val(r, "gpu__time_duration.sum") / 1e6 if "nsecond" in rows[1][col["gpu__time_duration.sum"]]
                   else val(r, "gpu__time_duration.sum"),
                   rd=val(r, "dram__bytes_read.sum"), wr=val(r, "dram__bytes_write.sum"),
                   rd_unit=rows[1][col["dram__bytes_read.sum"]], wr_unit=rows[1][col["dram__bytes_write.sum"]],
                   dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                   issue=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   fp64=val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                   warps=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                   regs=val(r, "launch__registers_per_thread"), inst=val(r, "smsp__inst_executed.sum"),
                   thr=val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {k: int(v["rd"] * scale.get(v["rd_unit"], 1.0) + v["wr"] * scale.get(v["wr_unit"], 1.0)) for k, v in full.items()}
json.dump({"source": f"profiles/{tag}_ncu_full_raw.csv (ncu --set full --clock-control none, one warm C2 step)",
           "workload": "C2 bright long burst x 14 detectors, one step", "dram_bytes_per_launch": traffic,
           "issue_slots_busy_frac": {k: round(v["issue"] / 100.0, 4) for k, v in full.items()}},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)

# launch list (ncu --metrics gpu__time_duration.sum): share of a step
lrows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 5]
lh = next(i for i, r in enumerate(lrows) if "Kernel Name" in r)
lc = {h: i for i, h in enumerate(lrows[lh])}
ncu_ms = {}
for r in lrows[lh + 1:]:
    try:
        v = float(r[lc["Metric Value"]].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = r[lc["Metric Unit"]]
    ms = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v if unit.startswith("ms") else v * 1e3
    n = short(r[lc["Kernel Name"]])
    ncu_ms[n] = ncu_ms.get(n, 0.0) + ms
b = json.loads(open(os.path.join(G, f"bench_{tag}.json")).read().strip().splitlines()[-1])
tot_ncu = sum(ncu_ms.values()) or 1.0
out = [f"# {title}", "",
       "Commands (each ncu pass after the same command exited 0 without ncu; `scripts/gpu_profile.sh " + tag + "`, artefacts by",
       "`scripts/make_profile.py`): `python bench.py --steps 5 --warmup 3` (the bench line), then",
       "`python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 0` under",
       "`ncu --metrics gpu__time_duration.sum --clock-control none` (launch list) and under",
       "`ncu --set full --clock-control none --import-source on` (full capture; `traffic.json` = DRAM bytes per launch and",
       "issue-slot utilisation, read by bench.py).", "",
       "## Bench line", "| | |", "|---|---|",
       f"| value (inputs resident) | {b['value']:.4g} {b['unit']}, {b['ms_per_step']:.3f} ms per burst, {b['gpu_launches']} launches in the timed region |",
       f"| e2e (`GBMGRB_CPL.go()`, H2D {b['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB + D2H {b['e2e']['d2h_bytes_per_step'] / 1e9:.2f} GB per burst inside the timed region) | {b['e2e']['value']:.4g} {b['e2e']['unit']} |"]
pop = b.get("population")
if pop:
    out.append(f"| population (`GBM_CPL_Universe.go(save=False, trigger=True)`, {pop['n_grbs']} bursts, trigger on the device) | "
               f"{pop['grbs_per_s']:.0f} GRB/s, {pop['events_per_s']:.4g} events/s in {pop['ms']:.1f} ms ({pop['kernel_ms_per_gpu']:.1f} ms of kernels), "
               f"{pop['gbm_trigger']['detected_local']} detected, {pop['n_skipped_per_gpu']} skipped |")
cb = b.get("cpu_baseline")
if cb:
    out.append(f"| cpu_baseline ({cb['kind']}, {cb['cores']} host threads) | {cb['value']:.4g} {cb['unit']} ({cb['sample']}) |")
out.append(f"| clocks | {b.get('clocks')} |")
out.append(f"| live yardsticks in the same run | {b.get('peaks_live')} |")
out += ["", "## Launch list: share of a step, ncu (cold, serialised) vs live CUDA events inside bench.py",
        "| kernel | ncu ms | ncu share | live ms | live share |", "|---|---|---|---|---|"]
for n, ms in sorted(ncu_ms.items(), key=lambda kv: -kv[1]):
    live = b["kernels"].get(n, {})
    out.append(f"| {n} | {ms:.3f} | {ms / tot_ncu:.3f} | {live.get('ms_per_step', float('nan')):.3f} | {live.get('share', float('nan')):.3f} |")
out += ["", "## `ncu --set full` per kernel",
        "| kernel | ms | DRAM read+write (MB) | DRAM % of peak | issue slots busy % | fp64 pipe % | warps active % | regs | warp inst | active thr / inst |",
        "|---|---|---|---|---|---|---|---|---|---|"]
for n, v in full.items():
    rd = v["rd"] * scale.get(v["rd_unit"], 1.0) / 1e6
    wr = v["wr"] * scale.get(v["wr_unit"], 1.0) / 1e6
    out.append(f"| {n} | {v['ms']:.3f} | {rd:.0f}+{wr:.0f} | {v['dram_pct']:.1f} | {v['issue']:.1f} | {v['fp64']:.1f} | {v['warps']:.1f} | "
               f"{v['regs']:.0f} | {v['inst']:.3g} | {v['thr']:.2f} |")
if pop:
    out += ["", "## Population kernels (ms per " + str(pop["n_grbs"]) + " bursts, live CUDA events)", "| kernel | ms |", "|---|---|"]
    out += [f"| {k} | {v:.3f} |" for k, v in pop["kernels"].items()]
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
