#!/usr/bin/env python
"""Where does the e2e arm's time go?  (a) the six-list copies of one bright burst alone (all kernels done before the
clock starts), in the engine's own copy pattern (one cudaMemcpyAsync per unit and list into pinned memory);
(b) GBMGRB_CPL.go() as bench.py times it; (c) one bare 1 GiB pinned copy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from cosmogrb_b200 import gbm, workloads
from cosmogrb_b200.engine import get_engine, Philox

eng = get_engine()
units, tabs = workloads.burst_inputs(dict(workloads.BRIGHT_LONG))
b = eng.new_batch(units, tabs)
eng.process(b, Philox(5))
torch.cuda.synchronize()
pinned = {}
for rep in range(4):
    t0 = time.perf_counter()
    out = eng.to_host(b, pinned)
    dt = time.perf_counter() - t0
    print(f"(a) copies alone, reused pinned buffers: {out['d2h_bytes'] / 1e9:.3f} GB in {1e3 * dt:.1f} ms = {out['d2h_bytes'] / dt / 1e9:.1f} GB/s")
for rep in range(3):
    t0 = time.perf_counter()
    out = eng.to_host(b, None)
    dt = time.perf_counter() - t0
    print(f"(a') copies alone, fresh pinned buffers: {1e3 * dt:.1f} ms = {out['d2h_bytes'] / dt / 1e9:.1f} GB/s")
    del out
grb = gbm.GBMGRB_CPL(ra=312.0, dec=-62.0, T0=0.1, name="probe", **dict(workloads.BRIGHT_LONG))
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    grb.go(seed=100 + rep)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"(b) go(): {1e3 * dt:.1f} ms")
buf = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
pin = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pin.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"(c) bare 1 GiB copy: {(1 << 30) / dt / 1e9:.1f} GB/s")
