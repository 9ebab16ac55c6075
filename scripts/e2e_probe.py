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

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
dist = None
if world > 1:           # all ranks of a box at once: they share the host side (torchrun --nproc-per-node N)
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def say(msg):
    print(f"[rank {rank}] {msg}", flush=True)


eng = get_engine()
units, tabs = workloads.burst_inputs(dict(workloads.BRIGHT_LONG))
b = eng.new_batch(units, tabs)
eng.process(b, Philox(5))
torch.cuda.synchronize()
pinned = {}
for rep in range(4):
    barrier()
    t0 = time.perf_counter()
    out = eng.to_host(b, pinned)
    dt = time.perf_counter() - t0
    say(f"(a) copies alone, reused pinned buffers: {out['d2h_bytes'] / 1e9:.3f} GB in {1e3 * dt:.1f} ms = {out['d2h_bytes'] / dt / 1e9:.1f} GB/s")
for rep in range(3):
    barrier()
    t0 = time.perf_counter()
    out = eng.to_host(b, None)
    dt = time.perf_counter() - t0
    say(f"(a') copies alone, fresh pinned buffers: {1e3 * dt:.1f} ms = {out['d2h_bytes'] / dt / 1e9:.1f} GB/s")
    del out
grb = gbm.GBMGRB_CPL(ra=312.0, dec=-62.0, T0=0.1, name="probe", **dict(workloads.BRIGHT_LONG))
for rep in range(4):
    barrier()
    t0 = time.perf_counter()
    grb.go(seed=100 + rep, stream_base=16 * rank)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    say(f"(b) go(): {1e3 * dt:.1f} ms")
buf = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
pin = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
for rep in range(3):
    barrier()
    t0 = time.perf_counter()
    pin.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    say(f"(c) bare 1 GiB copy: {(1 << 30) / dt / 1e9:.1f} GB/s")
barrier()
t0 = time.perf_counter()
for rep in range(4):        # sustained: 4 GiB back to back
    pin.copy_(buf, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
say(f"(c') bare 4 x 1 GiB back to back: {4 * (1 << 30) / dt / 1e9:.1f} GB/s")
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
