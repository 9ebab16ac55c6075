#!/usr/bin/env python
"""Turn one scripts/gpu_round2.sh run (gpurun_out/*_<tag>*) into the tracked artefacts under profiles/:
round2_<tag>_{bench.json, launches.csv, pop_launches.csv, c2_ncu_raw.csv, pop_ncu_raw.csv, c2_ncu_summary.txt,
pop_ncu_summary.txt, hot_lines.txt, summary.md} and traffic.json (DRAM bytes, issue-slot utilisation and warp
instructions per launch of the C2 step; DRAM bytes per event and issue-slot utilisation of the population kernels;
read by bench.py).  The .ncu-rep files themselves stay on the GPU box (gpurun returns at most 64 MiB): the raw and
source pages are exported there.
usage: scripts/make_profile2.py <tag> <pop_events> ["title line"]"""
import csv
import glob
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pop_events = float(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else f"round 2 profile {tag}"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
pre = os.path.join(P, f"round2_{tag}_")


def short(name):
    m = re.match(r"(?:void )?(?:cgrb::)?(\w+)(<[^>]*>)?", name)
    base, targs = m.group(1), m.group(2) or ""
    if base in ("k_gap_prefix", "k_seg_scan"):
        return base + targs.replace("(int)", "")
    if base == "k_level_order":
        return "k_background_order" if "1" in targs else "k_merge_order"
    return base


for src, dst in ((f"bench_{tag}.json", "bench.json"), (f"launches_{tag}.csv", "launches.csv"),
                 (f"pop_launches_{tag}.csv", "pop_launches.csv"), (f"c2_{tag}_raw.csv", "c2_ncu_raw.csv"),
                 (f"pop_{tag}_raw.csv", "pop_ncu_raw.csv"), (f"c2_{tag}_summary.txt", "c2_ncu_summary.txt"),
                 (f"pop_{tag}_summary.txt", "pop_ncu_summary.txt")):
    shutil.copy(os.path.join(G, src), pre + dst)
with open(pre + "hot_lines.txt", "w") as f:
    for path in sorted(glob.glob(os.path.join(G, f"c2_{tag}_hot_*.txt")) + glob.glob(os.path.join(G, f"pop_{tag}_hot_*.txt"))):
        f.write(f"==== {os.path.basename(path)} (ncu --page source: share of warp instructions / of stall samples per line)\n")
        f.write(open(path).read() + "\n")

SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TM = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}


def load_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")

    out = []
    for r in data:
        g = lambda k: val(r, k)                       # noqa: E731
        out.append(dict(
            name=short(r[col["Kernel Name"]]),
            ms=g("gpu__time_duration.sum") * TM.get(units[col["gpu__time_duration.sum"]], 1e-6),
            dram=g("dram__bytes_read.sum") * SC.get(units[col["dram__bytes_read.sum"]], 1.0)
            + g("dram__bytes_write.sum") * SC.get(units[col["dram__bytes_write.sum"]], 1.0),
            rd=g("dram__bytes_read.sum") * SC.get(units[col["dram__bytes_read.sum"]], 1.0),
            wr=g("dram__bytes_write.sum") * SC.get(units[col["dram__bytes_write.sum"]], 1.0),
            dram_pct=g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            issue=g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            fp64=g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            warps=g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            regs=g("launch__registers_per_thread"), inst=g("smsp__inst_executed.sum"), grid=g("launch__grid_size"),
            thr=g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            l1hit=g("l1tex__t_sector_hit_rate.pct"), l2hit=g("lts__t_sector_hit_rate.pct")))
    return out


c2 = load_raw(pre + "c2_ncu_raw.csv")
pop = load_raw(pre + "pop_ncu_raw.csv")


def launch_ms(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    c = {k: i for i, k in enumerate(rows[h])}
    tot, cnt = {}, {}
    for r in rows[h + 1:]:
        try:
            v = float(r[c["Metric Value"]].replace(",", ""))
        except (ValueError, KeyError):
            continue
        n = short(r[c["Kernel Name"]])
        tot[n] = tot.get(n, 0.0) + v * TM.get(r[c["Metric Unit"]], 1e-6)
        cnt[n] = cnt.get(n, 0) + 1
    return tot, cnt


b = json.loads(open(pre + "bench.json").read().strip().splitlines()[-1])
steps_in_list = 2      # the launch-list command runs --warmup 1 --steps 1 after new_batch
c2_ms, c2_cnt = launch_ms(pre + "launches.csv")
pop_ms, pop_cnt = launch_ms(pre + "pop_launches.csv")

popagg = {}
for k in pop:
    a = popagg.setdefault(k["name"], dict(ms=0.0, dram=0.0, inst=0.0, n=0, issue=0.0, warps=0.0, l1hit=0.0, l2hit=0.0, regs=k["regs"]))
    a["ms"] += k["ms"]
    a["dram"] += k["dram"]
    a["inst"] += k["inst"]
    a["n"] += 1
    for q in ("issue", "warps", "l1hit", "l2hit"):
        a[q] += k[q] * k["ms"]
traffic = {
    "source": f"profiles/round2_{tag}_c2_ncu_raw.csv and round2_{tag}_pop_ncu_raw.csv (ncu --set full --clock-control none; "
              "one warm C2 step; the first three batches of a 256-burst population, 10 GB of event buffers per batch)",
    "workload": "C2 bright long burst x 14 detectors, one step",
    "dram_bytes_per_launch": {k["name"]: int(k["dram"]) for k in c2},
    "issue_slots_busy_frac": {k["name"]: round(k["issue"] / 100.0, 4) for k in c2},
    "warp_inst_per_launch": {k["name"]: int(k["inst"]) for k in c2},
    "warps_active_frac": {k["name"]: round(k["warps"] / 100.0, 4) for k in c2},
    "population": {
        "workload": f"256 synthetic bursts (seed 12345), {pop_events:.0f} events, three batches; per-event figures divide by ALL events of the run",
        "dram_bytes_per_event": {n: round(a["dram"] / pop_events, 3) for n, a in popagg.items()},
        "warp_inst_per_event": {n: round(a["inst"] / pop_events, 4) for n, a in popagg.items()},
        "issue_slots_busy_frac": {n: round(a["issue"] / a["ms"] / 100.0, 4) for n, a in popagg.items()},
        "warps_active_frac": {n: round(a["warps"] / a["ms"] / 100.0, 4) for n, a in popagg.items()},
        "l1_hit_frac": {n: round(a["l1hit"] / a["ms"] / 100.0, 4) for n, a in popagg.items()},
        "l2_hit_frac": {n: round(a["l2hit"] / a["ms"] / 100.0, 4) for n, a in popagg.items()},
        "achieved_gbs_under_ncu": {n: round(a["dram"] / (a["ms"] * 1e-3) / 1e9, 1) for n, a in popagg.items()},
    },
}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)

out = [f"# {title}", "",
       f"`scripts/gpu_round2.sh {tag} 256` on one B200 through gpurun (every ncu pass after the same command exited 0 without ncu),",
       "artefacts by `scripts/make_profile2.py`.  Bench line: `python bench.py --steps 5 --warmup 3`.  C2 launch list and full",
       "capture: `python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population-total 0` under",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_` and `ncu --set full --clock-control none",
       "--import-source on` (the second, warm step).  Population: `python scripts/pop_run.py 256` (10 GB of event buffers per",
       "batch so that ncu's save/restore stays cheap), launch list + `--set full` of the hot kernels of the three batches.", "",
       "## Bench line", "| | |", "|---|---|",
       f"| value (inputs resident) | {b['value']:.4g} {b['unit']}, {b['ms_per_step']:.3f} ms per burst, {b['gpu_launches']} launches in the timed region |",
       f"| e2e (`GBMGRB_CPL.go()`, H2D {b['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB + D2H {b['e2e']['d2h_bytes_per_step'] / 1e9:.2f} GB per burst inside the timed region) | {b['e2e']['value']:.4g} {b['e2e']['unit']} "
       f"({b['e2e'].get('d2h_gbs_per_gpu', 0):.1f} GB/s device->host; bare-copy yardstick {b['e2e'].get('d2h_yardstick_gbs_per_gpu')} GB/s) |"]
q = b.get("population")
if q:
    out.append(f"| population C5 (`GBM_CPL_Universe.go(save=False, trigger=True)`, strong scaling) | {q['n_grbs']} bursts, {q['n_simulated']} simulated, "
               f"{q['n_skipped']} skipped: {q['grbs_per_s']:.0f} GRB/s, {q['events_per_s']:.4g} events/s in {q['ms'] / 1e3:.2f} s "
               + (f"({q['kernel_ms_per_gpu'] / 1e3:.2f} s of kernels, " if "kernel_ms_per_gpu" in q else
                  f"({q['kernel_ms_per_1000_bursts']:.1f} ms of kernels per 1000 bursts in the single-stream sample pass, ")
               + f"{q['batches_per_gpu']} batches) |")
cb = b.get("cpu_baseline")
if cb:
    out.append(f"| cpu_baseline ({cb['kind']}, {cb['cores']} host threads) | {cb['value']:.4g} {cb['unit']} ({cb['sample']}) |")
out.append(f"| clocks | {b.get('clocks')} |")
out.append(f"| live yardsticks in the same run | {b.get('peaks_live')} |")
tot_ncu = sum(c2_ms.values()) or 1.0
out += ["", "## C2 launch list: share of a step, ncu (cold cache, serialised) vs live CUDA events inside bench.py",
        "| kernel | launches in the command | ncu ms per launch | ncu share | live ms per step | live share |", "|---|---|---|---|---|---|"]
for n, ms in sorted(c2_ms.items(), key=lambda kv: -kv[1]):
    live = b["kernels"].get(n, {})
    out.append(f"| {n} | {c2_cnt[n]} | {ms / c2_cnt[n]:.3f} | {ms / tot_ncu:.3f} | {live.get('ms_per_step', float('nan')):.3f} | {live.get('share', float('nan')):.3f} |")
ev = b["config"]["events_per_step_per_gpu"]
out += ["", f"## C2 `ncu --set full`, one warm step ({ev} events, {b['config']['source_events']} photons)",
        "| kernel | ms | DRAM read+write (MB) | DRAM % of peak | issue slots busy % | fp64 pipe % | warps active % | regs | warp inst | active thr / inst | L1 hit % | L2 hit % |",
        "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for k in c2:
    out.append(f"| {k['name']} | {k['ms']:.3f} | {k['rd'] / 1e6:.0f}+{k['wr'] / 1e6:.0f} | {k['dram_pct']:.1f} | {k['issue']:.1f} | {k['fp64']:.1f} | "
               f"{k['warps']:.1f} | {k['regs']:.0f} | {k['inst']:.3g} | {k['thr']:.2f} | {k['l1hit']:.1f} | {k['l2hit']:.1f} |")
tot_dram = sum(k["dram"] for k in c2)
out += ["", f"Sum of DRAM bytes over the step: {tot_dram / 1e9:.2f} GB = {tot_dram / (18.0 * ev):.2f} x the algorithmic 18 B/event."]
e = next((k for k in c2 if k["name"] == "k_sample_energy_env"), None)
if e:
    out.append(f"`k_sample_energy_env`: {e['inst'] / b['config']['source_events']:.2f} warp instructions per photon "
               f"({b['config']['rejection_trials'] / b['config']['source_events']:.3f} trials per photon).")
if q:
    nsim = q["n_simulated"]
    if "kernels_ms_per_1000_bursts" in q:
        q = dict(q, kernels={k: v * nsim / 1000.0 for k, v in q["kernels_ms_per_1000_bursts"].items()})
        out += ["", q["kernel_sample"]]
    out += ["", f"## Population kernels: live CUDA events of the bench leg vs ncu on 256 bursts (ms per 1000 simulated bursts)",
            "| kernel | live | ncu launch list | ncu DRAM B/event | ncu GB/s | issue % | warps active % | L1 hit % | L2 hit % | regs |", "|---|---|---|---|---|---|---|---|---|---|"]
    for n, v in q["kernels"].items():
        a = popagg.get(n)
        nn = pop_ms.get(n, pop_ms.get(n + "<0>", float("nan")))
        row = f"| {n} | {v * 1000 / nsim:.2f} | {nn * 1000 / 255:.2f} |"
        if a:
            row += (f" {a['dram'] / pop_events:.1f} | {a['dram'] / (a['ms'] * 1e-3) / 1e9:.0f} | {a['issue'] / a['ms']:.1f} | {a['warps'] / a['ms']:.1f} | "
                    f"{a['l1hit'] / a['ms']:.1f} | {a['l2hit'] / a['ms']:.1f} | {a['regs']:.0f} |")
        else:
            row += " | | | | | | |"
        out.append(row)
open(pre + "summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
