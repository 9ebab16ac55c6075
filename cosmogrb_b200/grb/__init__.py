from ..utils.meta import RequiredParameter, SourceParameter
from .grb import GRB

__all__ = ["GRB", "SourceParameter", "RequiredParameter"]
