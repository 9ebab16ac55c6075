"""cosmogrb_b200 -- B200-native photon-event engine behind the cosmogrb API.

Only the per-photon hot path of grburgess/cosmogrb is implemented here (NHPP thinning, CPL
energy draws, DRM folding, GBM dead time, batched over GRB x detector units); see DESIGN.md.
The compute path is the CUDA library ``_lib/libcosmogrb_b200.so`` (C ABI in
``include/cosmogrb_b200.h``); there is no CPU fallback.
"""
from .config import cosmogrb_config

__version__ = "0.1.0"

__all__ = ["cosmogrb_config", "gbm", "GRBSave"]


def __getattr__(name):
    # lazy: keep `import cosmogrb_b200` light (no torch import) for the CPU-only checks
    if name == "gbm":
        import importlib

        return importlib.import_module(".instruments.gbm", __name__)
    if name == "GRBSave":
        from .io.grb_save import GRBSave

        return GRBSave
    raise AttributeError(name)
