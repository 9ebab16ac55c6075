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
    print(f"   roofline {r['kernel']} {r['bound']} ach {r['achieved']:.1f} / {r['peak']} = {r['frac']:.3f}")
    for k, v in d.get("kernels", {}).items():
        print(f"   {k:28s} {v['ms_per_step']:8.3f} ms  {v['share']:.3f}")
    for r in d.get("roofline_hbm_kernels", []):
        print(f"   hbm {r['kernel']:24s} {r['achieved']:.0f} GB/s frac {r['frac']:.3f}")
