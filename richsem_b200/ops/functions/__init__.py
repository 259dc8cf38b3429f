from .msda_function import MSDeformAttnFunction

__all__ = ["MSDeformAttnFunction"]
