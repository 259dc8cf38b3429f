from .msda_module import MSDeformAttn

__all__ = ["MSDeformAttn"]
