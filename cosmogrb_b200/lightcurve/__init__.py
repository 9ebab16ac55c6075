from .light_curve_storage import LightCurveStorage
from .lightcurve import LightCurve

__all__ = ["LightCurve", "LightCurveStorage"]
