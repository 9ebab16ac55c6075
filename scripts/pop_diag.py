#!/usr/bin/env python
"""Per-kernel live timings of a population run against batch size and run length, with the SM clock sampled
alongside: scripts/pop_diag.py <n_grbs> [batch_bytes]   (diagnostic for the ncu-vs-live gap of the streaming kernels)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
if len(sys.argv) > 2:
    os.environ["COSMOGRB_B200_BATCH_BYTES"] = sys.argv[2]
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
from bench import ClockSampler  # noqa: E402
from cosmogrb_b200 import _capi  # noqa: E402
from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe  # noqa: E402
from cosmogrb_b200.universe import synthetic_population  # noqa: E402

from cosmogrb_b200.engine import get_engine  # noqa: E402

_BATCH_SIZES = []
_eng = get_engine()
_orig_process = _eng.process


def _process(b, rng, **kw):
    _BATCH_SIZES.append(b.n_units // 14)
    return _orig_process(b, rng, **kw)


_eng.process = _process
warm = GBM_CPL_Universe(synthetic_population(min(n, 1000), seed=12345), save_path=None)
warm.go(save=False, seed=1, trigger=True)
uni = GBM_CPL_Universe(synthetic_population(n, seed=12345), save_path=None)
torch.cuda.synchronize()
cs = ClockSampler(0)
cs.start()
time.sleep(0.3)
_capi.profile_enable(True)
t0 = time.time()
uni.go(save=False, seed=1, trigger=True)
torch.cuda.synchronize()
t1 = time.time()
prof = _capi.profile_read()
_capi.profile_enable(False)
clk = cs.stop(t0, t1)
pk = {}
for name, m in prof:
    pk[name] = pk.get(name, 0.0) + m
nsim = n - len(uni.skipped)
print(f"n={n} simulated={nsim} batches={uni.n_batches} batch_bytes={os.environ.get('COSMOGRB_B200_BATCH_BYTES')} "
      f"wall {1e3 * (t1 - t0):.1f} ms = {nsim / (t1 - t0):.0f} GRB/s; kernels {sum(pk.values()):.1f} ms; clocks {clk}")
print(f"   host: {dict((k, round(v, 3)) for k, v in uni.timings.items())} {dict((k, round(v, 3)) for k, v in _eng.host_times.items())}")
print("   per 1000 bursts: " + ", ".join(f"{k} {v * 1000 / nsim:.2f}" for k, v in sorted(pk.items(), key=lambda kv: -kv[1])[:10]))
# per batch: events of each kind against the live time of the streaming kernels (is a slow kernel slow everywhere, or
# on the few batches that hold the bright bursts?)
if os.environ.get("POP_DIAG_BATCHES"):
    import numpy as np

    per, cur = [], None
    for name, m in prof:
        if name == "k_fmax":          # the sizing pass
            continue
        if cur is None:
            cur = {}
            per.append(cur)
        cur[name] = cur.get(name, 0.0) + m
        if name == "k_trig_eval":     # last launch of a batch
            cur = None
    ec = np.asarray(uni.event_counts)
    ok = ec[:, 0, 0] >= 0
    sizes = _BATCH_SIZES[-len(per):]
    order = np.nonzero(ok)[0]
    pos = 0
    for bi, (d, nb) in enumerate(zip(per, sizes)):
        g = order[pos:pos + nb]
        pos += nb
        ns, nk = int(ec[g, :, 0].sum()), int(ec[g, :, 1].sum())
        tot = ns + nk
        md = d.get("k_merge_deadtime", 0.0)
        print(f"   batch {bi:3d}: {nb:4d} bursts src {ns:.3e} bkg {nk:.3e} max-unit-src {int(ec[g, :, 0].max()):.2e} | merge {md:6.3f} ms "
              f"({18e-9 * tot / (md * 1e-3 + 1e-12):6.0f} GB/s) bkg {d.get('k_background', 0):6.3f} trig_bin {d.get('k_trig_bin', 0):6.3f} "
              f"trig_eval {d.get('k_trig_eval', 0):6.3f} energy {d.get('k_sample_energy_env', 0):6.3f} thin {d.get('k_source_thin', 0):6.3f}")
