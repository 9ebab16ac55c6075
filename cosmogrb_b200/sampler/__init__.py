from .background import Background, BackgroundSpectrumTemplate
from .constant_cpl import ConstantCPL
from .cpl_source import CPLSourceFunction
from .sampler import Sampler
from .source import Source
from .source_function import SourceFunction

__all__ = ["Sampler", "Source", "SourceFunction", "CPLSourceFunction", "ConstantCPL", "Background",
           "BackgroundSpectrumTemplate"]
