#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations that bench.py does not time (C1 README burst, C3 background-only
1000 s, C4 population sample) on one GPU, next to the CPU oracle (port of the reference's numba path) on all host
cores.  Prints one JSON object; results go into BASELINE.md §4.   scripts/config_bench.py > gpurun_out/configs.json"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
import torch  # noqa: E402

from cosmogrb_b200 import _capi  # noqa: E402
from cosmogrb_b200.engine import Philox, get_engine  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests import helpers as h  # noqa: E402

eng = get_engine()
dets = h.GBM_DETECTORS
tabs = [h.response(d).device_table(h.bkg_weights(d), "linear") for d in dets]
threads = len(os.sched_getaffinity(0))
out = {"host_threads": threads}


LAST_KERNELS = {}


def gpu_events_per_s(units, reps=20):
    b = eng.new_batch(units, tabs)
    for k in range(3):
        eng.reset_counts(b)
        eng.process(b, Philox(100 + k))
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps):
        eng.reset_counts(b)
        eng.process(b, Philox(200 + k))
    z.record()
    z.synchronize()
    c = eng.fetch_counts(b)
    eng.check_status(b)
    ev = int(c["n_src"].sum() + c["n_bkg"].sum())
    ms = a.elapsed_time(z) / reps
    # per-kernel durations of one more step (CUDA events around every launch)
    _capi.profile_enable(True)
    eng.reset_counts(b)
    eng.process(b, Philox(999))
    torch.cuda.synchronize()
    LAST_KERNELS.clear()
    for name, m in _capi.profile_read():
        LAST_KERNELS[name] = round(LAST_KERNELS.get(name, 0.0) + m, 4)
    _capi.profile_enable(False)
    return ev, ms, ev / (ms * 1e-3)


def cpu_events_per_s(params, bkg, source=True):
    """the oracle's LightCurve.process() for the 14 detectors, one detector per thread"""
    def one(i):
        rsp = h.response(dets[i])
        s = h.oracle_source(rsp, **params)
        n, _ = orc.process_detector(s, rsp.energy_edges, rsp.effective_area_packed, rsp.cumulative_matrix,
                                    h.bkg_weights(dets[i]), 0.0, params["duration"] if source else 0.0, bkg[0], bkg[1],
                                    bkg[2], seed_numba=11 + i, seed_numpy=97 + i)
        return n
    one(0)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        n = sum(ex.map(one, range(len(dets))))
    dt = time.perf_counter() - t0
    return n, dt, n / dt


# C1: README burst, 14 detectors
p1 = h.README_DEMO
u1 = np.concatenate([h.unit(det_index=i, stream_id=i, **p1) for i in range(14)])
ev, ms, rate = gpu_events_per_s(u1)
n, dt, crate = cpu_events_per_s(p1, (-100.0, 300.0, 500.0))
out["C1"] = {"gpu_events": ev, "gpu_ms_per_burst": ms, "gpu_events_per_s": rate, "gpu_grbs_per_s": 1e3 / ms,
             "cpu_events": n, "cpu_s_per_burst": dt, "cpu_events_per_s": crate, "cpu_grbs_per_s": 1.0 / dt,
             "gpu_kernels_ms": dict(LAST_KERNELS), "gpu_kernels_ms_sum": round(sum(LAST_KERNELS.values()), 4)}

# C3: background only, 1000 s, 14 detectors
p3 = dict(h.README_DEMO, peak_flux=5e-20)
u3 = np.concatenate([h.unit(det_index=i, stream_id=50 + i, flags=_capi.UNIT_NO_SOURCE, bkg=(0.0, 1000.0, 500.0), **p3)
                     for i in range(14)])
ev, ms, rate = gpu_events_per_s(u3)
n, dt, crate = cpu_events_per_s(p3, (0.0, 1000.0, 500.0), source=False)
out["C3"] = {"gpu_events": ev, "gpu_ms": ms, "gpu_events_per_s": rate, "cpu_events": n, "cpu_s": dt,
             "cpu_events_per_s": crate}

# C4 on the CPU: the first 16 ordinary bursts of the synthetic population (runaway ones excluded), one burst = 14 detectors
from cosmogrb_b200.universe import synthetic_population  # noqa: E402

pop = synthetic_population(1000, seed=12345)
picked, t_cpu, n_cpu = 0, 0.0, 0
for g in range(1000):
    if picked >= 16:
        break
    p = dict(z=pop.distances[g], peak_flux=pop.fluxes.latent[g], alpha=pop.alpha[g], ep_start=pop.ep[g],
             ep_tau=pop.tau[g], trise=pop.trise[g], tdecay=pop.tdecay[g], duration=pop.duration[g])
    rsp = h.response("n0")
    fm = orc.fmax(h.oracle_source(rsp, **p), rsp.effective_area_packed, 0.0, p["duration"])
    if not np.isfinite(fm) or fm * p["duration"] > 3e5:      # keep the sample cheap and representative of the bulk
        continue
    n, dt, _ = cpu_events_per_s(p, (-100.0, 300.0, 500.0))
    picked, t_cpu, n_cpu = picked + 1, t_cpu + dt, n_cpu + n
out["C4_cpu"] = {"bursts": picked, "events": n_cpu, "seconds": t_cpu, "grbs_per_s": picked / t_cpu,
                 "events_per_s": n_cpu / t_cpu,
                 "note": "bursts with fmax*T <= 3e5 candidates per detector (the bulk); 14 detectors on the host threads"}
print(json.dumps(out))
