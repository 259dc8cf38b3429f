"""ctypes binding of libmsda_b200.so (C ABI: include/msda_b200.h).

The library is the product; there is no Python or PyTorch fallback.  If it is missing or does not
export the expected ABI, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from pathlib import Path

import torch

# MSDA_B200_LIB selects another build of the same library (kernel-tuning experiments only)
_LIB_PATH = Path(os.environ.get("MSDA_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libmsda_b200.so")

MSDA_OK = 0
MSDA_ERR_UNSUPPORTED = 2
MSDA_ABI_VERSION = 3
FLAG_DETERMINISTIC = 0x1
FLAG_GRAD_VALUE_PREZEROED = 0x2
FLAG_FORCE_GENERIC = 0x4
FLAG_COORDS_FMA = 0x10
FLAG_NO_GRAD_VALUE = 0x400
# msda_opts.kernel_hint (testing / tuning: which kernel family serves a head_dim-32 problem; 0 = the library decides)
KERNEL_AUTO, KERNEL_SPLIT, KERNEL_TILED, KERNEL_WINDOW = 0, 1, 2, 3
MAX_LEVELS = 16

_vp, _i, _i64p = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)


class MsdaOpts(ctypes.Structure):
    """struct msda_opts (include/msda_b200.h)."""

    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("flags", ctypes.c_uint32),
        ("spatial_shapes_host", _i64p),
        ("level_start_index_host", _i64p),
        ("query_order", _vp),
        ("query_order_len", ctypes.c_int32),
        ("kernel_hint", ctypes.c_int32),
        ("workspace", _vp),
        ("workspace_bytes", ctypes.c_size_t),
    ]


def _load():
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} not found: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m richsem_b200._build`) "
            "from the repo root. richsem_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(str(_LIB_PATH))
    lib.msda_abi_version.restype = _i
    if lib.msda_abi_version() != MSDA_ABI_VERSION:
        raise RuntimeError(f"{_LIB_PATH}: ABI {lib.msda_abi_version()} != expected {MSDA_ABI_VERSION}; rebuild")
    opts_p = ctypes.POINTER(MsdaOpts)
    fwd = [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, opts_p]
    bwd = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, opts_p]
    for sfx in ("f32", "f64", "bf16"):
        f = getattr(lib, f"msda_forward_{sfx}")
        f.argtypes, f.restype = fwd, _i
        b = getattr(lib, f"msda_backward_{sfx}")
        b.argtypes, b.restype = bwd, _i
    fwd_fused = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, opts_p]
    bwd_fused = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, opts_p]
    for sfx in ("f32", "bf16"):
        f = getattr(lib, f"msda_forward_fused_{sfx}")
        f.argtypes, f.restype = fwd_fused, _i
        b = getattr(lib, f"msda_backward_fused_{sfx}")
        b.argtypes, b.restype = bwd_fused, _i
    _ll = ctypes.c_longlong
    lib.msda_value_prepare_bf16.argtypes = [_vp, _vp, _vp, _ll, _i, _vp]
    lib.msda_value_prepare_bf16.restype = _i
    lib.msda_zero_masked_rows_f32.argtypes = [_vp, _vp, _vp, _ll, _i]
    lib.msda_zero_masked_rows_f32.restype = _i
    lib.msda_encoder_proposals_f32.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, opts_p]
    lib.msda_encoder_proposals_f32.restype = _i
    lib.msda_encoder_proposals_backward_f32.argtypes = [_vp, _vp, _vp, _i, _i, _i, _vp]
    lib.msda_encoder_proposals_backward_f32.restype = _i
    lib.msda_rowmax_f32.argtypes = [_vp, _vp, _ll, _i, _vp]
    lib.msda_rowmax_f32.restype = _i
    lib.msda_topk_rows_f32.argtypes = [_vp, _vp, _i, _i, _i, _vp, _vp]
    lib.msda_topk_rows_f32.restype = _i
    _f = ctypes.c_float
    lib.msda_add_layernorm_f32.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _vp, _vp, _vp]
    lib.msda_add_layernorm_f32.restype = _i
    lib.msda_add_layernorm_backward_f32.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp, _vp, _vp, _vp,
                                                    ctypes.c_size_t]
    lib.msda_add_layernorm_backward_f32.restype = _i
    lib.msda_add_layernorm_workspace_bytes.argtypes = [_ll, _i]
    lib.msda_add_layernorm_workspace_bytes.restype = ctypes.c_size_t
    lib.msda_debug_corners_f32.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, opts_p]
    lib.msda_debug_corners_f32.restype = _i
    lib.msda_backward_workspace_bytes.argtypes = [_i] * 7
    lib.msda_backward_workspace_bytes.restype = ctypes.c_size_t
    lib.msda_has_fast_path.argtypes = [_i] * 4
    lib.msda_has_fast_path.restype = _i
    lib.msda_build_info.restype = ctypes.c_char_p
    lib.msda_last_error.restype = ctypes.c_char_p
    lib.msda_launch_count.restype = ctypes.c_uint64
    return lib


lib = _load()
LIB_PATH = str(_LIB_PATH)


def check(status: int, what: str) -> None:
    if status != MSDA_OK:
        raise RuntimeError(f"{what} failed (status {status}): {lib.msda_last_error().decode()}")


def build_info() -> str:
    return lib.msda_build_info().decode()


def launch_count() -> int:
    return int(lib.msda_launch_count())


# ----------------------------------------------------------------------------------------
# Host mirror of the level table.
#
# spatial_shapes / level_start_index arrive as device int64 tensors (reference:
# ms_deform_attn_cuda.cu:67-68) but the kernels want them in constant memory, so the host needs
# their values.  Reading them is a device->host sync; it is done once per tensor object and
# cached for as long as that object is alive and unmodified (id + _version + weakref), which is
# safe against allocator address reuse.
# ----------------------------------------------------------------------------------------
class LevelMeta:
    __slots__ = ("shapes", "starts", "c_shapes", "c_starts", "spatial_size_sum")

    def __init__(self, shapes, starts):
        self.shapes = tuple((int(h), int(w)) for h, w in shapes)
        self.starts = tuple(int(s) for s in starts)
        flat = [x for hw in self.shapes for x in hw]
        self.c_shapes = (ctypes.c_int64 * len(flat))(*flat)
        self.c_starts = (ctypes.c_int64 * len(self.starts))(*self.starts)
        self.spatial_size_sum = sum(h * w for h, w in self.shapes)


_meta_cache: dict = {}
_meta_lock = threading.Lock()


def _tensor_key(t: torch.Tensor):
    # inference tensors (created under torch.inference_mode) do not track a version counter and cannot be
    # modified in place outside inference mode: identity alone keys them
    return (id(t), -1 if t.is_inference() else t._version)


def level_meta(spatial_shapes: torch.Tensor, level_start_index: torch.Tensor) -> LevelMeta:
    key = (_tensor_key(spatial_shapes), _tensor_key(level_start_index))
    hit = _meta_cache.get(key)
    if hit is not None:
        return hit[0]
    if spatial_shapes.dim() != 2 or spatial_shapes.shape[1] != 2:
        raise RuntimeError(f"spatial_shapes must be (num_levels, 2), got {tuple(spatial_shapes.shape)}")
    if level_start_index.numel() != spatial_shapes.shape[0]:
        raise RuntimeError("level_start_index must have one entry per level")
    meta = LevelMeta(spatial_shapes.tolist(), level_start_index.tolist())  # the one host sync
    with _meta_lock:
        if len(_meta_cache) > 256:
            _meta_cache.clear()

        def _drop(_ref, key=key):
            _meta_cache.pop(key, None)

        try:
            refs = (weakref.ref(spatial_shapes, _drop), weakref.ref(level_start_index, _drop))
        except TypeError:
            return meta
        _meta_cache[key] = (meta, refs)
    return meta


# ----------------------------------------------------------------------------------------
# Query processing order (cache-locality hint, never affects results)
# ----------------------------------------------------------------------------------------
_order_cache: dict = {}
PATCH = 8  # queries are tiled in PATCH x PATCH pixel patches per level; 64 = kTileQ of the kernels


def build_patch_order(shapes, starts, pad=False):
    """Order of the S encoder tokens that visits each level in 8x8 pixel patches (row-major inside a
    patch).  A thread block takes 64 consecutive entries.
    pad=False: a permutation of [0, S); edge patches are smaller, so later blocks straddle two
               neighbouring patches.
    pad=True:  every patch occupies exactly 64 entries, missing pixels are -1 (skipped by the
               kernels), so that one block = one patch everywhere (tighter windows for the L1 and
               for the on-chip grad_value merging of the window backward, at the price of some idle lane groups)."""
    import numpy as np

    parts = []
    for (h, w), st in zip(shapes, starts):
        if not pad:
            ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            key = ((ys // PATCH) * ((w + PATCH - 1) // PATCH) + xs // PATCH) * (PATCH * PATCH) \
                + (ys % PATCH) * PATCH + xs % PATCH
            parts.append(st + np.argsort(key.ravel(), kind="stable"))
        else:
            sup = int(os.environ.get("MSDA_B200_SUPER_PATCH", "1"))  # experiment: SUPxSUP patches per super-patch
            big = PATCH * sup
            hp, wp = (h + big - 1) // big * big, (w + big - 1) // big * big
            ys, xs = np.meshgrid(np.arange(hp), np.arange(wp), indexing="ij")
            tok = np.where((ys < h) & (xs < w), st + ys * w + xs, -1)
            # (super row, sub row, y, super col, sub col, x) -> super-patch major, then 8x8 sub-patch, then pixel
            tok = tok.reshape(hp // big, sup, PATCH, wp // big, sup, PATCH).transpose(0, 3, 1, 4, 2, 5).reshape(-1)
            parts.append(tok)
    return np.concatenate(parts).astype(np.int32)


def query_order(meta: LevelMeta, num_query: int, device) -> "torch.Tensor | None":
    """Patch-tiled order for encoder self-attention (num_query == spatial size, levels laid out
    back to back); None (natural order) otherwise."""
    mode = os.environ.get("MSDA_B200_QUERY_ORDER", "patch")
    if mode == "natural" or num_query != meta.spatial_size_sum:
        return None
    acc = 0
    for (h, w), st in zip(meta.shapes, meta.starts):
        if st != acc:
            return None
        acc += h * w
    pad = os.environ.get("MSDA_B200_ORDER_PAD", "1") not in ("", "0")
    key = (meta.shapes, str(device), pad, os.environ.get("MSDA_B200_SUPER_PATCH", "1"))
    t = _order_cache.get(key)
    if t is None:
        t = torch.from_numpy(build_patch_order(meta.shapes, meta.starts, pad=pad)).to(device)
        if len(_order_cache) > 64:
            _order_cache.clear()
        _order_cache[key] = t
    return t


_opts_cache: dict = {}


def make_opts(meta: LevelMeta, order=None, flags=0, workspace=None, kernel=KERNEL_AUTO) -> MsdaOpts:
    """msda_opts for a launch.  Structs are immutable once built, so they are cached per
    (level table, order buffer, flags, workspace, kernel hint) to keep the per-call host cost down."""
    key = (id(meta), order.data_ptr() if order is not None else 0, flags,
           workspace.data_ptr() if workspace is not None else 0,
           workspace.numel() if workspace is not None else 0, kernel)
    hit = _opts_cache.get(key)
    if hit is not None and hit[1] is meta:
        return hit[0]
    o = _build_opts(meta, order, flags, workspace, kernel)
    if len(_opts_cache) > 512:
        _opts_cache.clear()
    _opts_cache[key] = (o, meta, order, workspace)  # keep the referenced buffers alive
    return o


def _build_opts(meta, order, flags, workspace, kernel=KERNEL_AUTO) -> MsdaOpts:
    o = MsdaOpts()
    o.struct_size = ctypes.sizeof(MsdaOpts)
    o.flags = flags
    o.kernel_hint = kernel
    o.spatial_shapes_host = ctypes.cast(meta.c_shapes, _i64p)
    o.level_start_index_host = ctypes.cast(meta.c_starts, _i64p)
    if order is not None:
        o.query_order = order.data_ptr()
        o.query_order_len = order.numel()
    if workspace is not None:
        o.workspace = workspace.data_ptr()
        o.workspace_bytes = workspace.numel() * workspace.element_size()
    return o
