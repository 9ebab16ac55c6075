"""Fermi-GBM instrument layer (mirror of cosmogrb/instruments/gbm/__init__.py)."""
import importlib

_LAZY = {
    "GBMGRB": ".gbm_grb", "GBMGRB_CPL": ".gbm_grb", "GBMGRB_CPL_Constant": ".gbm_grb",
    "GBMLightCurve": ".gbm_lightcurve", "GBMBackground": ".gbm_background",
    "NaIResponse": ".synthetic_response", "BGOResponse": ".synthetic_response",
    "GBM_CPL_Universe": ".gbm_universe", "GBM_CPL_Constant_Universe": ".gbm_universe",
    "GBMLightCurveAnalyzer": ".gbm_lightcurve_analyzer", "GBMTrigger": ".gbm_trigger",
}

__all__ = sorted(_LAZY)


def __getattr__(name):
    if name in _LAZY:
        return getattr(importlib.import_module(_LAZY[name], __name__), name)
    raise AttributeError(name)
