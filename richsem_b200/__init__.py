"""richsem_b200 — B200-native (sm_100a) multi-scale deformable attention, a drop-in for the one
data-parallel hot path of MengLcool/RichSem (``models/richsem/ops``).

    from richsem_b200.ops.modules import MSDeformAttn
    from richsem_b200.ops.functions import MSDeformAttnFunction
    import richsem_b200.MultiScaleDeformableAttention as MSDA   # the compiled extension's surface

Importing the package loads richsem_b200/lib/libmsda_b200.so and fails loudly if it is missing:
there is no CPU or PyTorch fallback.
"""
from . import _capi  # noqa: F401  (loads the CUDA library or raises)
from . import MultiScaleDeformableAttention  # noqa: F401
from .MultiScaleDeformableAttention import is_deterministic, set_coords_fma, set_deterministic  # noqa: F401

__version__ = "0.1.0"


def install_as_reference_extension() -> None:
    """Register this package's extension surface under the reference's import name, so that the
    reference's own ``functions/ms_deform_attn_func.py`` (``import MultiScaleDeformableAttention as
    MSDA``) runs on these kernels unmodified."""
    import sys

    sys.modules["MultiScaleDeformableAttention"] = MultiScaleDeformableAttention
