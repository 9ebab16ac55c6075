"""``GBMBackground`` (cosmogrb/instruments/gbm/gbm_background.py:7-50): the per-detector 128-channel
background template.  The templates are the reference's own data (data/gbm_backgrounds/<det>.h5),
shipped here decoded in data/gbm_tables.npz (scripts/extract_gbm_tables.py)."""
from ...sampler.background import Background, BackgroundSpectrumTemplate
from .synthetic_response import GBM_DETECTORS, gbm_tables

_allowed_gbm_detectors = GBM_DETECTORS

_templates = {}


def gbm_background_template(detector):
    if detector not in _templates:
        counts = gbm_tables()["bkg_counts"][GBM_DETECTORS.index(detector)]
        _templates[detector] = BackgroundSpectrumTemplate(counts, start_at_one=False)
    return _templates[detector]


class GBMBackground(Background):
    def __init__(self, tstart, tstop, average_rate=1000, detector=None):
        assert detector in _allowed_gbm_detectors, "must use a proper GBM detector n<> or b<>"
        super(GBMBackground, self).__init__(tstart=tstart, tstop=tstop, average_rate=average_rate,
                                            background_spectrum_template=gbm_background_template(detector))
