from .response import Response

__all__ = ["Response"]
