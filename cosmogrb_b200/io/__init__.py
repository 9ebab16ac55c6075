from .gbm_fits import grbsave_to_gbm_fits
from .grb_save import GRBSave

__all__ = ["GRBSave", "grbsave_to_gbm_fits"]
