"""``GBMLightCurve`` (cosmogrb/instruments/gbm/gbm_lightcurve.py:11-115): a LightCurve whose merged
event list goes through the GBM dead-time filter on the GPU (literal reference semantics by default;
``cosmogrb_config.engine.deadtime_mode = "physical"`` selects the non-paralyzable model)."""
import numpy as np

from ...engine import get_engine
from ...lightcurve.lightcurve import LightCurve


class GBMLightCurve(LightCurve):
    _deadtime = True

    def __init__(self, source, background, response, name, grb_name, tstart, tstop):
        super(GBMLightCurve, self).__init__(source=source, background=background, response=response, name=name,
                                            grb_name=grb_name, tstart=tstart, tstop=tstop, instrument="GBM")
        self._extra_info["angle"] = self._response.separation_angle          # gbm_lightcurve.py:40

    def _filter_deadtime(self):
        """stand-alone filter on ``self._times`` / ``self._pha`` (gbm_lightcurve.py:42-56)."""
        sel = get_engine().gbm_dead_time(self._times, self._pha)
        self._times = np.asarray(self._times)[sel]
        self._pha = np.asarray(self._pha)[sel]

    def write_tte(self, file_name=None):
        """TTE FITS file of the (dead-time-filtered) event list, ``<grb>_<detector>.fits``
        (gbm_lightcurve.py:58-77); needs ``process()`` to have run."""
        from ...utils.tte_file import TTEFile

        if self._times is None or self._pha is None:
            raise RuntimeError("write_tte() needs a processed light curve: call process() first")
        rsp = self._response
        tte_file = TTEFile(
            det_name=rsp.detector_name,
            tstart=self._background.tstart + rsp.T0,
            tstop=self._background.tstop + rsp.T0,
            trigger_time=rsp.trigger_time,
            ra=rsp.ra,
            dec=rsp.dec,
            channel=np.arange(0, 128, dtype=np.int16),
            emin=rsp.channel_edges[:-1],
            emax=rsp.channel_edges[1:],
            pha=np.asarray(self._pha).astype(np.int16),
            time=np.asarray(self._times, dtype="f8") + rsp.T0,
        )
        file_name = file_name or f"{self._grb_name}_{self._name}.fits"
        tte_file.writeto(file_name, overwrite=True)
        return file_name


def gbm_dead_time(time, pha, mode=None):
    """``_gbm_dead_time`` (gbm_lightcurve.py:80-115) on the GPU: returns (filtered_time, filtered_pha,
    selection) with the reference's conventions (zeros where an event is dropped).  ``mode``: None (the
    engine's configured model) | "reference" | "physical"."""
    time, pha = np.asarray(time, dtype="f8"), np.asarray(pha, dtype="f8")
    sel = get_engine().gbm_dead_time(time, pha, mode=mode)
    return np.where(sel, time, 0.0), np.where(sel, pha, 0.0), sel.astype("f8")
