"""Import-path alias: the reference keeps MSDeformAttn in modules/ms_deform_attn.py."""
from .msda_module import MSDeformAttn  # noqa: F401
