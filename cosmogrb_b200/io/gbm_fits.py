"""``grbsave_to_gbm_fits`` (cosmogrb/io/gbm_fits.py:9-83): a saved burst -> one TTE file and one RSP file
per detector (``tte_<grb>_<det>.fits``, ``rsp_<grb>_<det>.rsp``), the products 3ML / the GBM tools read."""
import os

import numpy as np

from ..utils.tte_file import TTEFile
from .grb_save import GRBSave


def grbsave_to_gbm_fits(file_name, destination=".", detectors=None):
    grb = file_name if isinstance(file_name, GRBSave) else GRBSave.from_file(file_name)
    if detectors is None:
        detectors = list(grb.keys())
    else:
        for det in detectors:
            assert det in grb, f"{det} is not a in the light curve list"
    files = {}
    for key, det in grb.items():
        if key not in detectors:
            continue
        lc, rsp = det["lightcurve"], det["response"]
        shift = lc.time_adjustment
        tte_file_name = os.path.join(destination, f"tte_{grb.name}_{key}.fits")
        TTEFile(det_name=key, tstart=lc.tstart + shift, tstop=lc.tstop + shift, trigger_time=grb.T0 + shift,
                ra=grb.ra, dec=grb.dec, channel=np.arange(0, 128, dtype=np.int16), emin=rsp.channel_edges[:-1],
                emax=rsp.channel_edges[1:], pha=np.asarray(lc.pha).astype(np.int16),
                time=np.asarray(lc.times) + shift).writeto(tte_file_name, overwrite=True)
        rsp_file_name = os.path.join(destination, f"rsp_{grb.name}_{key}.rsp")
        rsp.to_fits(rsp_file_name, telescope_name="fermi", instrument_name=key, overwrite=True)
        files[key] = dict(tte=tte_file_name, rsp=rsp_file_name)
    return files


__all__ = ["grbsave_to_gbm_fits"]
