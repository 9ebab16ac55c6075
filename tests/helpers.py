"""Shared test inputs: synthetic responses, unit records (from the package's workload definitions), oracle source
structs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cosmogrb_b200 import _capi  # noqa: E402,F401
from cosmogrb_b200.instruments.gbm.synthetic_response import (  # noqa: E402,F401
    GBM_DETECTORS, BGOResponse, NaIResponse, gbm_tables)
from cosmogrb_b200.workloads import (BRIGHT_LONG, CONFTEST_BRIGHT, README_DEMO, bkg_weights,  # noqa: E402,F401
                                     burst_inputs, response, unit)


def oracle_source(rsp, kind=0, interp_extrap="linear", **p):
    from oracle import oracle as orc

    return orc.source(kind=kind, interp_extrap=interp_extrap, peak_flux=p["peak_flux"],
                      ep_start=p.get("ep", p.get("ep_start")), ep_tau=p.get("ep_tau", 1.0), alpha=p["alpha"],
                      trise=p.get("trise", 1.0), tdecay=p.get("tdecay", 1.0), z=p["z"], emin=rsp.emin,
                      emax=rsp.emax)
