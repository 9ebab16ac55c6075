"""GBM bursts: 14-detector fan-out (cosmogrb/instruments/gbm/gbm_grb.py:18-157).

``_setup`` builds one (response, background, source, light curve) per detector exactly like the
reference; the DRMs come from the synthetic GBM-shaped generator (gbm_drm_gen / gbmgeometry are not
installable offline).  The 14 ``Source`` objects need fmax: it is computed for all detectors in one
device launch instead of 14 x 50 host calls.
"""
import numpy as np

from ...engine import get_engine
from ...grb import GRB, SourceParameter
from ...sampler.constant_cpl import ConstantCPL
from ...sampler.cpl_source import CPLSourceFunction
from ...sampler.source import Source
from ...utils.logging import setup_logger
from .gbm_background import GBMBackground
from .gbm_lightcurve import GBMLightCurve
from .synthetic_response import GBM_DETECTORS, BGOResponse, GBMResponseFile, NaIResponse

logger = setup_logger(__name__)


class _LazyFmaxSource(Source):
    """Source whose fmax is filled in by the owning GRB for all detectors at once."""

    def _get_energy_integrated_max(self):
        return None


class GBMGRB(GRB):
    _background_start = -100
    _background_stop = 300
    _gbm_detectors = GBM_DETECTORS

    def __init__(self, source_function_class, response_files=None, **kwargs):
        """``response_files`` (not in the reference, whose DRMs come from ``gbm_drm_gen``): ``{detector: path}``
        of OGIP ``.rsp`` files (140 x 128) to use instead of the synthetic matrices for those detectors."""
        self._use_plaw_sample = True
        self._response_files = dict(response_files or {})
        for det in self._response_files:
            assert det in self._gbm_detectors, f"{det} is not a GBM detector"
        super(GBMGRB, self).__init__(source_function_class=source_function_class, **kwargs)

    def _setup_source(self):
        sources = []
        for key in self._responses.keys():
            source_function = self._source_function(response=self._responses[key], **self._source_params)
            source = _LazyFmaxSource(0.0, self.duration, source_function, z=self.z,
                                     use_plaw_sample=self._use_plaw_sample)
            sources.append(source)
            lc = GBMLightCurve(source, self._backgrounds[key], self._responses[key], name=key, grb_name=self.name,
                               tstart=self._background_start, tstop=self._background_stop)
            lc.set_time_adjustment(self._responses[key].trigger_time)
            self._add_lightcurve(lc)
        # Source._get_energy_integrated_max for the 14 detectors in one launch (source.py:88-110)
        eng = get_engine()
        lcs = list(self._lightcurves.values())
        tabs = [lc.device_table(eng.interp_extrap) for lc in lcs]
        units = np.concatenate([lc.unit_record(det_index=i) for i, lc in enumerate(lcs)])
        b = eng.new_batch(units, tabs, compute_fmax=True)
        for s, f in zip(sources, b.units["fmax"]):
            s._fmax = float(f)

    def _setup(self):
        for det in self._gbm_detectors:
            if det in self._response_files:
                rsp = GBMResponseFile(det, self.ra, self.dec, self._response_files[det], name=self.name)
            else:
                cls = BGOResponse if det[0] == "b" else NaIResponse
                rsp = cls(det, self.ra, self.dec, save=False, name=self.name)
            self._add_response(det, rsp)
            bkg = GBMBackground(self._background_start, self._background_stop, average_rate=500, detector=det)
            self._add_background(det, bkg)
        self._setup_source()


class GBMGRB_CPL(GBMGRB):
    peak_flux = SourceParameter()
    alpha = SourceParameter()
    ep_start = SourceParameter()
    ep_tau = SourceParameter()
    trise = SourceParameter()
    tdecay = SourceParameter()

    def __init__(self, response_files=None, **kwargs):
        super(GBMGRB_CPL, self).__init__(source_function_class=CPLSourceFunction, response_files=response_files, **kwargs)


class GBMGRB_CPL_Constant(GBMGRB):
    peak_flux = SourceParameter()
    alpha = SourceParameter()
    ep = SourceParameter()

    def __init__(self, response_files=None, **kwargs):
        super(GBMGRB_CPL_Constant, self).__init__(source_function_class=ConstantCPL, response_files=response_files, **kwargs)


__all__ = ["GBMGRB", "GBMGRB_CPL", "GBMGRB_CPL_Constant"]
