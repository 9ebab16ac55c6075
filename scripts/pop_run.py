#!/usr/bin/env python
"""A small population run for profiling: scripts/pop_run.py [n_grbs]  (GBM trigger on the device)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe  # noqa: E402
from cosmogrb_b200.universe import synthetic_population  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
uni = GBM_CPL_Universe(synthetic_population(n, seed=12345), save_path=None)
uni.go(save=False, seed=1, trigger=True)
print(f"{n} GRBs, {int(uni.detected.sum())} detected, {len(uni.skipped)} skipped, "
      f"{int(uni.event_counts[uni.event_counts[:, 0, 0] >= 0][:, :, :2].sum())} events")
