"""Shared test inputs: synthetic responses, unit records, oracle source structs."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cosmogrb_b200 import _capi  # noqa: E402
from cosmogrb_b200.instruments.gbm.synthetic_response import (  # noqa: E402
    GBM_DETECTORS, BGOResponse, NaIResponse, gbm_tables)

README_DEMO = dict(z=1.0, peak_flux=5e-7, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=1.0, tdecay=1.0,
                   duration=80.0)
CONFTEST_BRIGHT = dict(z=1.0, peak_flux=1e-5, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=0.1, tdecay=1.0,
                       duration=2.0)
BRIGHT_LONG = dict(z=1.0, peak_flux=1e-4, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=5.0, tdecay=60.0,
                   duration=300.0)


@functools.lru_cache(maxsize=None)
def response(det="n0", ra=312.0, dec=-62.0):
    cls = BGOResponse if det[0] == "b" else NaIResponse
    return cls(det, ra, dec)


def bkg_weights(det="n0"):
    t = gbm_tables()
    c = t["bkg_counts"][GBM_DETECTORS.index(det)]
    return c / c.sum()        # background.py:69-73 (float32 arithmetic like the reference)


def unit(kind=0, det_index=0, stream_id=0, flags=0, bkg=(-100.0, 300.0, 500.0), **p):
    u = np.zeros(1, dtype=_capi.UNIT_DTYPE)
    u["ea_scale"] = 1.0
    u["kind"], u["det"], u["stream_id"], u["flags"] = kind, det_index, stream_id, flags
    u["peak_flux"], u["alpha"], u["z"] = p["peak_flux"], p["alpha"], p["z"]
    u["ep_start"] = p.get("ep", p.get("ep_start"))
    u["ep_tau"], u["trise"], u["tdecay"] = p.get("ep_tau", 1.0), p.get("trise", 1.0), p.get("tdecay", 1.0)
    u["src_tstart"], u["src_tstop"] = 0.0, p["duration"]
    u["bkg_tstart"], u["bkg_tstop"], u["bkg_rate"] = bkg
    return u


def oracle_source(rsp, kind=0, interp_extrap="linear", **p):
    from oracle import oracle as orc      # lazy: bench.py's GPU arm imports this module without touching oracle/

    return orc.source(kind=kind, interp_extrap=interp_extrap, peak_flux=p["peak_flux"],
                      ep_start=p.get("ep", p.get("ep_start")), ep_tau=p.get("ep_tau", 1.0), alpha=p["alpha"],
                      trise=p.get("trise", 1.0), tdecay=p.get("tdecay", 1.0), z=p["z"], emin=rsp.emin,
                      emax=rsp.emax)
