from .population import Population, synthetic_population
from .universe import GRBWrapper, ParameterServer, Universe

__all__ = ["Universe", "ParameterServer", "GRBWrapper", "Population", "synthetic_population"]
