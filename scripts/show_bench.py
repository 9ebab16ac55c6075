#!/usr/bin/env python
"""Print the key numbers of a bench.py JSON line: scripts/show_bench.py gpurun_out/bench_X.json"""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    c = d["config"]
    print(f"== {path}: value {d['value']:.4g} {d['unit']}  {d['ms_per_step']:.3f} ms/step  n_gpus={d['n_gpus']}  launches={d.get('gpu_launches')}")
    print(f"   e2e {d['e2e']['value']:.4g}  cpu {d.get('cpu_baseline') and d['cpu_baseline']['value']}  clocks {d.get('clocks')}")
    print(f"   events {c.get('events_per_step_per_gpu')} cand {c.get('candidates')} trials {c.get('rejection_trials')} exact {c.get('exact_lambda_evaluations')}")
    r = d["roofline"]
    print(f"   roofline {r['kernel']} {r['bound']} ach {r['achieved']} / {r['peak']} = {r['frac']}  warp inst / event {r.get('warp_inst_per_event')}")
    for k, v in d.get("kernels", {}).items():
        print(f"   {k:28s} {v['ms_per_step']:8.3f} ms  {v['share']:.3f}")
    for r in d.get("roofline_hbm_kernels", []):
        print(f"   hbm {r['kernel']:24s} {r['achieved']:.0f} GB/s frac {r['frac']:.3f}")
    if d.get("population"):
        q = d["population"]
        print(f"   population: {q['n_grbs']} GRBs ({q.get('n_simulated')} simulated) {q['events']:.4g} events in {q['ms']:.1f} ms -> {q['grbs_per_s']:.1f} GRB/s "
              f"{q['events_per_s']:.4g} ev/s; kernels {q.get('kernel_ms_per_gpu', q.get('kernel_ms_per_1000_bursts', 0)):.1f} ms"
              f"{' per 1000 bursts (sample pass)' if 'kernel_ms_per_1000_bursts' in q else ''}, batches {q['batches_per_gpu']}")
        print(f"      host seconds: {q.get('host_seconds')}  detail: {q.get('host_detail')}")
        for k, v in q.get("kernels", q.get("kernels_ms_per_1000_bursts", {})).items():
            print(f"      {k:26s} {v:9.3f} ms")
        for k, v in q["hbm_kernels"].items():
            print(f"      hbm {k:22s} {v['achieved_gbs']:.0f} GB/s frac {v['frac_of_measured_hbm']:.3f}")
