"""Named configurations of BASELINE.json and the synthetic inputs they run on: burst parameters,
GBM-shaped synthetic responses, background templates and unit records.  Used by bench.py, scripts/ and the tests."""
import functools

import numpy as np

from . import _capi
from .instruments.gbm.synthetic_response import GBM_DETECTORS, BGOResponse, NaIResponse, gbm_tables

# C1: the README demo burst (README.md:37-49; ep -> ep_start, tau -> ep_tau)
README_DEMO = dict(z=1.0, peak_flux=5e-7, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=1.0, tdecay=1.0,
                   duration=80.0)
# the reference's bright test fixture (test/conftest.py:33-45)
CONFTEST_BRIGHT = dict(z=1.0, peak_flux=1e-5, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=0.1, tdecay=1.0,
                       duration=2.0)
# C2: bright long burst; trise / tdecay chosen so that the 50-point fmax grid sees the pulse
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
    """one (GRB, detector) unit record from burst parameters"""
    u = np.zeros(1, dtype=_capi.UNIT_DTYPE)
    u["ea_scale"] = 1.0
    u["kind"], u["det"], u["stream_id"], u["flags"] = kind, det_index, stream_id, flags
    u["peak_flux"], u["alpha"], u["z"] = p["peak_flux"], p["alpha"], p["z"]
    u["ep_start"] = p.get("ep", p.get("ep_start"))
    u["ep_tau"], u["trise"], u["tdecay"] = p.get("ep_tau", 1.0), p.get("trise", 1.0), p.get("tdecay", 1.0)
    u["src_tstart"], u["src_tstop"] = 0.0, p["duration"]
    u["bkg_tstart"], u["bkg_tstop"], u["bkg_rate"] = bkg
    return u


def burst_inputs(params, grb_index=0, interp_extrap="linear", bkg=(-100.0, 300.0, 500.0)):
    """Host-side inputs of one burst seen by the 14 GBM detectors: 14 unit records + 14 detector table blocks."""
    tabs = [response(d).device_table(bkg_weights(d), interp_extrap) for d in GBM_DETECTORS]
    units = np.concatenate([unit(det_index=i, stream_id=16 * grb_index + i, bkg=bkg, **params)
                            for i in range(len(GBM_DETECTORS))])
    return units, tabs
