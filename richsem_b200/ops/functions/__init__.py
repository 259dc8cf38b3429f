from .msda_function import MSDeformAttnFunction, MSDeformAttnFusedFunction

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFusedFunction"]
