#!/usr/bin/env python
"""GPU timeline of a population run: per batch, when its first kernel could start and when its last one ended
(CUDA events on the compute stream), against the host clock.  scripts/pop_timeline.py [n_grbs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cosmogrb_b200.engine import get_engine  # noqa: E402
from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe  # noqa: E402
from cosmogrb_b200.universe import synthetic_population  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
eng = get_engine()
for rep in range(2):
    uni = GBM_CPL_Universe(synthetic_population(n, seed=12345), save_path=None)
    marks = []
    orig_process, orig_trig = eng.process, eng.gbm_trigger

    def process(b, rng, **kw):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        marks.append([time.perf_counter(), e0, None, None])
        return orig_process(b, rng, **kw)

    def trig(b, *a, **kw):
        r = orig_trig(b, *a, **kw)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        marks[-1][2], marks[-1][3] = e1, time.perf_counter()
        return r

    eng.process, eng.gbm_trigger = process, trig
    torch.cuda.synchronize()
    p0 = torch.cuda.Event(enable_timing=True)
    p0.record()
    h0 = time.perf_counter()
    uni.go(save=False, seed=1, trigger=True)
    p1 = torch.cuda.Event(enable_timing=True)
    p1.record()
    torch.cuda.synchronize()
    h1 = time.perf_counter()
    eng.process, eng.gbm_trigger = orig_process, orig_trig
    if rep == 0:
        continue
    print(f"total: gpu {p0.elapsed_time(p1):.1f} ms, host {1e3 * (h1 - h0):.1f} ms; host timings {uni.timings}")
    prev_end = 0.0
    for i, (th0, e0, e1, th1) in enumerate(marks):
        s, e = p0.elapsed_time(e0), p0.elapsed_time(e1)
        print(f"batch {i}: host enqueue {1e3 * (th0 - h0):7.1f} -> {1e3 * (th1 - h0):7.1f} ms | gpu start {s:7.1f} end {e:7.1f} "
              f"busy {e - s:6.1f} idle-before {s - prev_end:6.1f}")
        prev_end = e
