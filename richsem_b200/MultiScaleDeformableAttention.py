"""Drop-in for RichSem's compiled extension module ``MultiScaleDeformableAttention``.

Same two entry points, argument order and return types as the pybind11 module of the reference
(/root/reference/models/richsem/ops/src/vision.cpp:13-16, prototypes ms_deform_attn.h:20-27,41-49,
host wrappers cuda/ms_deform_attn_cuda.cu:20-80 and :83-153), implemented on top of the C ABI of
libmsda_b200.so.  To use it under the reference's import name::

    import sys, richsem_b200.MultiScaleDeformableAttention as m
    sys.modules["MultiScaleDeformableAttention"] = m

Error behaviour mirrors the reference: precondition failures raise RuntimeError (the reference's
AT_ASSERTM); CPU tensors raise "Not implemented on the CPU" (ms_deform_attn.h:38).  Kernel launch
failures also raise (the reference only printf()s them).

Beyond the reference: bfloat16 ``value`` (fp32 locations / weights, bf16 output; ``ms_deform_attn_backward`` returns
fp32 gradients — the autograd Function then casts grad_value to the bf16 input's dtype, as autograd requires), a
deterministic grad_value mode (``set_deterministic``) and the choice of pixel-coordinate arithmetic
(``set_coords_fma``).
"""
from __future__ import annotations

import os

import torch

from . import _capi

_SUFFIX = {torch.float32: "f32", torch.float64: "f64", torch.bfloat16: "bf16"}
_FWD = {k: getattr(_capi.lib, "msda_forward_" + k) for k in ("f32", "f64", "bf16")}
_BWD = {k: getattr(_capi.lib, "msda_backward_" + k) for k in ("f32", "f64", "bf16")}
_state = {"deterministic": os.environ.get("MSDA_B200_DETERMINISTIC", "0") not in ("", "0"),
          "coords_fma": os.environ.get("MSDA_B200_COORDS_FMA", "0") not in ("", "0")}
_workspaces: dict = {}


def set_deterministic(flag: bool) -> None:
    """grad_value by sort-by-corner segmented sums (bitwise reproducible) instead of fp32 atomics."""
    _state["deterministic"] = bool(flag)


def set_coords_fma(flag: bool) -> None:
    """Pixel coordinate = fma(loc, size, -0.5) instead of the default round(loc * size) - 0.5.

    The default is the reference SOURCE (ms_deform_im2col_cuda.cuh:285-286, a rounded multiply followed by a rounded
    subtract) and is what every parity claim of this repo refers to.  nvcc compiles that source line with -fmad=true
    into ONE fused multiply-add, so the reference's shipped binary computes the fused form; the two differ only where
    loc * size rounds onto the pixel lattice — every sample of an un-jittered initialisation (zero offset weights,
    integer-pixel offset bias), where the floor cell flips and grad_sampling_loc takes the other one-sided derivative.
    Set this (or MSDA_B200_COORDS_FMA=1) to reproduce the compiled reference sample for sample."""
    _state["coords_fma"] = bool(flag)


def is_deterministic() -> bool:
    return _state["deterministic"] or torch.are_deterministic_algorithms_enabled()


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _check_inputs(named):
    for name, t in named:
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    for name, t in named:
        if not t.is_cuda:
            if name == "value":
                raise RuntimeError("Not implemented on the CPU")
            if name not in ("spatial_shapes", "level_start_index"):
                raise RuntimeError(f"{name} must be a CUDA tensor")


def _dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step):
    _require(value.dim() == 4, "value must be (N, S, M, D)")
    _require(sampling_loc.dim() == 6 and sampling_loc.shape[-1] == 2, "sampling_loc must be (N, Lq, M, L, P, 2)")
    batch, spatial_size, num_heads, channels = value.shape
    num_levels = spatial_shapes.shape[0]
    num_query, num_point = sampling_loc.shape[1], sampling_loc.shape[4]
    _require(tuple(sampling_loc.shape) == (batch, num_query, num_heads, num_levels, num_point, 2),
             "sampling_loc shape does not match value / spatial_shapes")
    _require(tuple(attn_weight.shape) == (batch, num_query, num_heads, num_levels, num_point),
             "attn_weight shape does not match sampling_loc")
    # The reference splits the batch into chunks of min(batch, im2col_step) images and launches once
    # per chunk (ms_deform_attn_cuda.cu:50-75); images are independent, so one launch covers them
    # all here, but the divisibility requirement is kept so that callers see the same errors.
    step = min(batch, int(im2col_step)) if batch > 0 else 1
    _require(step > 0 and batch % step == 0, f"batch({batch}) must divide im2col_step({step})")
    return batch, spatial_size, num_heads, channels, num_levels, num_query, num_point


def _aux_dtype(value):
    return torch.float32 if value.dtype == torch.bfloat16 else value.dtype


def _suffix(value, sampling_loc, attn_weight):
    sfx = _SUFFIX.get(value.dtype)
    if sfx is None:
        raise RuntimeError(f'"ms_deform_attn" not implemented for \'{value.dtype}\'')
    aux = _aux_dtype(value)
    _require(sampling_loc.dtype == aux and attn_weight.dtype == aux,
             f"sampling_loc / attn_weight must be {aux} when value is {value.dtype}")
    return sfx


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device):
    if _raw_stream is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


def _dev_ptr(t):
    return t.data_ptr() if t.is_cuda else None


_get_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


class _on_device:
    """Makes `device` current for the launch; free when it already is (the common case)."""

    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if _get_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


# --------------------------------------------------------------------------------------------
# Per-call host path.  A decoder layer's kernels take 14-36 us, so the Python in front of a launch matters: the
# first call with a given (level tensors, shapes, dtype, device) validates everything the reference's host wrappers
# validate (ms_deform_attn_cuda.cu:28-52, 93-105) and records a _Plan; later calls with the same signature only
# re-check what can change between calls (contiguity, the weights' shape, dtypes, in-place edits of the level
# tensors) and go straight to the allocation and the C call.
# --------------------------------------------------------------------------------------------
class _Plan:
    __slots__ = ("shp", "st", "shp_ver", "st_ver", "wshape", "aux", "sfx", "dims", "out_shape", "meta", "order",
                 "shp_ptr", "st_ptr", "opts", "go_numel")


_plans: dict = {}  # tests that change MSDA_B200_QUERY_ORDER between calls clear it


def _ver(t):
    return -1 if t.is_inference() else t._version


def _lookup_plan(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    try:
        key = (id(spatial_shapes), id(level_start_index), value.shape, sampling_loc.shape, value.dtype, value.device,
               im2col_step)
        p = _plans.get(key)
    except (AttributeError, TypeError):
        return None, None
    if (p is not None and p.shp() is spatial_shapes and p.st() is level_start_index and p.shp_ver == _ver(spatial_shapes)
            and p.st_ver == _ver(level_start_index) and attn_weight.shape == p.wshape and sampling_loc.dtype == p.aux
            and attn_weight.dtype == p.aux and sampling_loc.is_cuda and attn_weight.is_cuda and value.is_contiguous()
            and sampling_loc.is_contiguous() and attn_weight.is_contiguous()):
        return p, key
    return None, key


def _make_plan(key, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """Full validation (the reference's checks and messages), then the cached plan."""
    import weakref

    _check_inputs((("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)))
    dims = _dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step)
    p = _Plan()
    p.sfx = _suffix(value, sampling_loc, attn_weight)
    p.dims = dims
    n, s, m, d, nl, lq, npt = dims
    p.meta = _capi.level_meta(spatial_shapes, level_start_index)
    p.order = _capi.query_order(p.meta, lq, value.device)
    p.wshape, p.aux = attn_weight.shape, _aux_dtype(value)
    p.out_shape, p.go_numel = (n, lq, m * d), n * lq * m * d
    p.shp_ptr, p.st_ptr = _dev_ptr(spatial_shapes), _dev_ptr(level_start_index)
    p.shp_ver, p.st_ver = _ver(spatial_shapes), _ver(level_start_index)
    p.opts = {}
    if key is not None and value.is_cuda:
        try:
            p.shp, p.st = weakref.ref(spatial_shapes), weakref.ref(level_start_index)
            if len(_plans) > 256:
                _plans.clear()
            _plans[key] = p
        except TypeError:
            pass
    return p


def _plan_opts(p, flags, kernel, ws):
    k = (flags, kernel, ws.data_ptr() if ws is not None else 0, ws.numel() if ws is not None else 0)
    o = p.opts.get(k)
    if o is None:
        if len(p.opts) > 64:
            p.opts.clear()
        o = p.opts[k] = (_capi._build_opts(p.meta, p.order, flags, ws, kernel), ws)  # keeps the workspace alive
    return o[0]


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step,
                           _flags: int = 0, _kernel: int = 0):
    p, key = _lookup_plan(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    if p is None:
        p = _make_plan(key, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    out = torch.empty(p.out_shape, dtype=value.dtype, device=value.device)
    if out.numel() == 0:
        return out
    if _state["coords_fma"]:
        _flags |= _capi.FLAG_COORDS_FMA
    n, s, m, d, nl, lq, npt = p.dims
    with _on_device(value.device):
        rc = _FWD[p.sfx](
            _stream(value.device), value.data_ptr(), p.shp_ptr, p.st_ptr, sampling_loc.data_ptr(), attn_weight.data_ptr(),
            n, s, m, d, nl, lq, npt, out.data_ptr(), _plan_opts(p, _flags, _kernel, None))
    if rc:
        _capi.check(rc, "msda_forward_" + p.sfx)
    return out


def _workspace(nbytes, device, stream):
    # one workspace per (device, stream): two backward calls in flight on different streams must not share the
    # fixed-point accumulators
    key = (device.type, device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if len(_workspaces) > 16:
            _workspaces.clear()
        ws = torch.empty(int(nbytes * 1.1) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step, _flags: int = 0, _need_grad_value: bool = True, _kernel: int = 0):
    """Same arguments and return value as the reference's binding (vision.cpp:15).  `_need_grad_value=False`
    (an addition; the autograd Function passes `ctx.needs_input_grad[0]`) skips the grad_value scatter — the
    more expensive half of the backward — and returns None in its place."""
    p, key = _lookup_plan(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    if p is None:
        _require(grad_output.is_contiguous(), "grad_output tensor has to be contiguous")
        _require(grad_output.is_cuda or not value.is_cuda, "grad_output must be a CUDA tensor")
        p = _make_plan(key, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    _require(grad_output.is_contiguous(), "grad_output tensor has to be contiguous")
    _require(grad_output.is_cuda, "grad_output must be a CUDA tensor")
    _require(grad_output.dtype == value.dtype and grad_output.numel() == p.go_numel,
             "grad_output must have value's dtype and shape (N, Lq, M*D)")
    # zero-filled by the library
    grad_value = torch.empty(value.shape, dtype=p.aux, device=value.device) if _need_grad_value else None
    grad_loc = torch.empty_like(sampling_loc)
    grad_attw = torch.empty_like(attn_weight)
    if grad_loc.numel() == 0:  # no queries (or empty batch): nothing is sampled
        return [grad_value.zero_() if _need_grad_value else None, grad_loc, grad_attw]
    n, s, m, d, nl, lq, npt = p.dims
    flags, ws = _flags, None
    if _state["coords_fma"]:
        flags |= _capi.FLAG_COORDS_FMA
    if not _need_grad_value:
        flags |= _capi.FLAG_NO_GRAD_VALUE
    if (_state["deterministic"] or torch.are_deterministic_algorithms_enabled()) and value.dtype != torch.float64:
        flags |= _capi.FLAG_DETERMINISTIC
    stream = _stream(value.device)
    if flags & _capi.FLAG_DETERMINISTIC:
        ws = _workspace(_capi.lib.msda_backward_workspace_bytes(n, s, m, d, nl, lq, npt), value.device, stream)
    with _on_device(value.device):
        rc = _BWD[p.sfx](
            stream, grad_output.data_ptr(), value.data_ptr(), p.shp_ptr, p.st_ptr, sampling_loc.data_ptr(),
            attn_weight.data_ptr(), n, s, m, d, nl, lq, npt, grad_value.data_ptr() if _need_grad_value else None,
            grad_loc.data_ptr(), grad_attw.data_ptr(), _plan_opts(p, flags, _kernel, ws))
    if rc:
        _capi.check(rc, "msda_backward_" + p.sfx)
    return [grad_value, grad_loc, grad_attw]


# --------------------------------------------------------------------------------------------
# Fused prologue: the module's softmax and sampling-location arithmetic
# (/root/reference/models/richsem/ops/modules/ms_deform_attn.py:98-111) done inside the kernels.
# --------------------------------------------------------------------------------------------
_FWD_FUSED = {k: getattr(_capi.lib, "msda_forward_fused_" + k) for k in ("f32", "bf16")}
_BWD_FUSED = {k: getattr(_capi.lib, "msda_backward_fused_" + k) for k in ("f32", "bf16")}


def fused_prologue_supported(value, num_levels, num_query, num_point) -> bool:
    """True when the library has fused-prologue kernels for this problem: CUDA fp32 / bf16 value (16-byte aligned),
    head_dim 32, 4 points, 3 to 5 levels, default (non-deterministic) mode, and a problem served by the split kernels
    (at most 65,536 (query, head) pairs: decoder cross-attention) or the window kernels (encoder self-attention:
    num_query == spatial size, for which the host builds the patch order)."""
    if not (value.is_cuda and value.dtype in (torch.float32, torch.bfloat16) and value.dim() == 4 and value.shape[3] == 32
            and num_point == 4 and 3 <= num_levels <= 5 and value.data_ptr() % 16 == 0 and not is_deterministic()):
        return False
    pairs = value.shape[0] * num_query * value.shape[2]
    return 0 < pairs <= 65536 or (num_query == value.shape[1] and os.environ.get("MSDA_B200_QUERY_ORDER", "patch") != "natural")


def _fused_args(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits, im2col_step):
    named = (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("reference_points", reference_points), ("sampling_offsets", sampling_offsets), ("attn_logits", attn_logits))
    _check_inputs(named)
    _require(reference_points.dim() == 4 and reference_points.shape[-1] in (2, 4),
             "reference_points must be (N, Lq, L, 2) or (N, Lq, L, 4)")
    n, lq, nl, ref_dim = reference_points.shape
    _require(sampling_offsets.dim() == 6, "sampling_offsets must be (N, Lq, M, L, P, 2)")
    npt = sampling_offsets.shape[4]
    dims = _dims(value, spatial_shapes, sampling_offsets, attn_logits.view(n, lq, value.shape[2], nl, npt), im2col_step)
    _require(reference_points.dtype == torch.float32 and sampling_offsets.dtype == torch.float32
             and attn_logits.dtype == torch.float32, "reference_points / sampling_offsets / attn_logits must be float32")
    return dims, ref_dim


def ms_deform_attn_forward_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                 attn_logits, im2col_step, _flags: int = 0, _kernel: int = 0):
    """out (N, Lq, M*D) from the RAW sampling offsets (N, Lq, M, L, P, 2) and attention logits (N, Lq, M, L*P)."""
    (n, s, m, d, nl, lq, npt), ref_dim = _fused_args(value, spatial_shapes, level_start_index, reference_points,
                                                     sampling_offsets, attn_logits, im2col_step)
    sfx = _SUFFIX[value.dtype]
    meta = _capi.level_meta(spatial_shapes, level_start_index)
    out = torch.empty((n, lq, m * d), dtype=value.dtype, device=value.device)
    opts = _capi.make_opts(meta, order=_capi.query_order(meta, lq, value.device), flags=_flags, kernel=_kernel)
    with _on_device(value.device):
        rc = _FWD_FUSED[sfx](
            _stream(value.device), value.data_ptr(), _dev_ptr(spatial_shapes), _dev_ptr(level_start_index),
            sampling_offsets.data_ptr(), attn_logits.data_ptr(), reference_points.data_ptr(), ref_dim, n, s, m, d, nl, lq,
            npt, out.data_ptr(), opts)
    _capi.check(rc, "msda_forward_fused_" + sfx)
    return out


def ms_deform_attn_backward_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                  attn_logits, grad_output, im2col_step, _flags: int = 0, _kernel: int = 0):
    """[grad_value, grad_sampling_offsets, grad_attn_logits]; no gradient for reference_points."""
    (n, s, m, d, nl, lq, npt), ref_dim = _fused_args(value, spatial_shapes, level_start_index, reference_points,
                                                     sampling_offsets, attn_logits, im2col_step)
    sfx = _SUFFIX[value.dtype]
    _require(grad_output.is_contiguous() and grad_output.dtype == value.dtype and grad_output.numel() == n * lq * m * d,
             "grad_output must be contiguous, with value's dtype and shape (N, Lq, M*D)")
    meta = _capi.level_meta(spatial_shapes, level_start_index)
    grad_value = torch.empty(value.shape, dtype=torch.float32, device=value.device)  # zero-filled by the library
    grad_offsets = torch.empty_like(sampling_offsets)
    grad_logits = torch.empty_like(attn_logits)
    opts = _capi.make_opts(meta, order=_capi.query_order(meta, lq, value.device), flags=_flags, kernel=_kernel)
    with _on_device(value.device):
        rc = _BWD_FUSED[sfx](
            _stream(value.device), grad_output.data_ptr(), value.data_ptr(), _dev_ptr(spatial_shapes),
            _dev_ptr(level_start_index), sampling_offsets.data_ptr(), attn_logits.data_ptr(),
            reference_points.data_ptr(), ref_dim, n, s, m, d, nl, lq, npt, grad_value.data_ptr(),
            grad_offsets.data_ptr(), grad_logits.data_ptr(), opts)
    _capi.check(rc, "msda_backward_fused_" + sfx)
    return [grad_value, grad_offsets, grad_logits]


def debug_corners(spatial_shapes, level_start_index, sampling_loc, _flags: int = 0):
    """int32 (N,Lq,M,L,P,4) bilinear corner token indices as the fp32/bf16 kernels compute them
    (-1 = contributes nothing).  Test hook for the bit-exact index contract."""
    _require(sampling_loc.is_cuda and sampling_loc.dtype == torch.float32 and sampling_loc.is_contiguous(),
             "sampling_loc must be a contiguous CUDA float32 tensor")
    n, lq, m, nl, npt, _ = sampling_loc.shape
    meta = _capi.level_meta(spatial_shapes, level_start_index)
    out = torch.empty((n, lq, m, nl, npt, 4), dtype=torch.int32, device=sampling_loc.device)
    with _on_device(sampling_loc.device):
        rc = _capi.lib.msda_debug_corners_f32(_stream(sampling_loc.device), _dev_ptr(spatial_shapes), _dev_ptr(level_start_index),
                                              sampling_loc.data_ptr(), n, m, nl, lq, npt, out.data_ptr(),
                                              _capi.make_opts(meta, flags=_flags))
    _capi.check(rc, "msda_debug_corners_f32")
    return out
