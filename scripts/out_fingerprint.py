#!/usr/bin/env python
"""Fingerprint of the production path's output (refactors that must not change a single event):
    [COSMOGRB_B200_LIB=variant.so] python scripts/out_fingerprint.py
sha256 over the six event lists of every detector of one bright burst (CONFTEST_BRIGHT at 10x its flux, ~3e6 photons),
the trial counters, and the event counts + trigger results of a 300-burst population.  Two builds whose kernels draw the
same Philox counters and do the same arithmetic print the same line."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cosmogrb_b200 import workloads as h  # noqa: E402
from cosmogrb_b200.engine import Engine, Philox  # noqa: E402
from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe  # noqa: E402
from cosmogrb_b200.universe import synthetic_population  # noqa: E402

eng = Engine(device=torch.device("cuda", 0))
params = dict(h.CONFTEST_BRIGHT, peak_flux=1e-4)
units, tabs = h.burst_inputs(params, grb_index=3)
b = eng.new_batch(units, tabs)
eng.process(b, Philox(7))
c = eng.fetch_counts(b)
eng.check_status(b)
m = hashlib.sha256()
for d in range(len(units)):
    for name, kind in (("times_src", "src"), ("pha_src", "src"), ("times_bkg", "bkg"), ("pha_bkg", "bkg"),
                       ("times_out", "kept"), ("pha_out", "kept")):
        m.update(np.ascontiguousarray(b.view(name, d, kind)).tobytes())
for k in ("n_src", "n_bkg", "n_kept", "n_trials", "n_candidates"):
    m.update(np.ascontiguousarray(c[k]).tobytes())
n_ev = int(c["n_src"].sum() + c["n_bkg"].sum())
uni = GBM_CPL_Universe(synthetic_population(300, seed=4321), save_path=None)
uni.go(save=False, seed=5, trigger=True)
m.update(np.ascontiguousarray(uni.event_counts).tobytes())
m.update(np.ascontiguousarray(uni.detected).tobytes())
print(f"fingerprint {m.hexdigest()[:32]}  burst events {n_ev} trials {int(c['n_trials'].sum())}  "
      f"population events {int(uni.event_counts[uni.event_counts[:, 0, 0] >= 0][:, :, :2].sum())} "
      f"detected {int(uni.detected.sum())}  lib {os.environ.get('COSMOGRB_B200_LIB', 'in-tree')}")
