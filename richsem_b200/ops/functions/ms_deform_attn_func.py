"""Import-path alias: the reference keeps MSDeformAttnFunction in functions/ms_deform_attn_func.py."""
from .msda_function import MSDeformAttnFunction  # noqa: F401
