from .grb_save import GRBSave

__all__ = ["GRBSave"]
