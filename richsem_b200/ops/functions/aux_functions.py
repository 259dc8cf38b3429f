"""Host side of the elementwise passes either side of the sampling core (SURVEY section 8f rows 2 and 4;
C ABI: include/msda_b200.h, kernels: csrc/msda_aux.cu).

* ``prepare_value`` — the module's padding-mask zeroing of the projected value
  (/root/reference/models/richsem/ops/modules/ms_deform_attn.py:94-97), fused with the bf16 cast the bf16
  kernels want.  fp32: the masked rows are zeroed in place (nothing else is touched).
* ``gen_encoder_output_proposals`` — same name, arguments and results as the reference helper
  (/root/reference/models/richsem/utils.py:10-65), one kernel pass instead of ~25 PyTorch kernels.

* ``add_layer_norm`` — ``norm(src + src2)``, the encoder layer's epilogue around ``output_proj`` and around the FFN
  (/root/reference/models/richsem/deformable_transformer.py:871-872, 866-867), one pass forward and one backward.

CUDA only, like everything in this package: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _capi
from ...MultiScaleDeformableAttention import _on_device, _require, _stream


def _mask_bytes(mask, rows_shape):
    _require(mask.dtype in (torch.bool, torch.uint8), "padding mask must be bool (or uint8)")
    _require(tuple(mask.shape) == tuple(rows_shape), f"padding mask must be {tuple(rows_shape)}, got {tuple(mask.shape)}")
    _require(mask.is_cuda, "padding mask must be a CUDA tensor")
    mask = mask.contiguous()
    return mask.view(torch.uint8) if mask.dtype == torch.bool else mask


_shapes_cache: dict = {}


def _shapes_host(spatial_shapes):
    """ctypes int64 mirror of ``spatial_shapes``; a device tensor is read once per (object, version)."""
    if not isinstance(spatial_shapes, torch.Tensor):
        flat = [int(x) for hw in spatial_shapes for x in hw]
        return (ctypes.c_int64 * len(flat))(*flat)
    key = (id(spatial_shapes), -1 if spatial_shapes.is_inference() else spatial_shapes._version)
    hit = _shapes_cache.get(key)
    if hit is not None and hit[1]() is spatial_shapes:
        return hit[0]
    flat = [int(x) for hw in spatial_shapes.tolist() for x in hw]  # the one host sync
    arr = (ctypes.c_int64 * len(flat))(*flat)
    if len(_shapes_cache) > 64:
        _shapes_cache.clear()
    try:
        _shapes_cache[key] = (arr, weakref.ref(spatial_shapes))
    except TypeError:
        pass
    return arr


def zero_masked_rows_(data: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """In place: ``data[..., :] = 0`` on the rows where ``mask`` is set (no autograd)."""
    if not data.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    _require(data.dtype == torch.float32 and data.is_contiguous(), "data must be a contiguous fp32 tensor")
    m = _mask_bytes(mask, data.shape[:-1])
    rows = m.numel()
    with _on_device(data.device):
        _capi.check(_capi.lib.msda_zero_masked_rows_f32(_stream(data.device), data.data_ptr(), m.data_ptr(), rows,
                                                        data.shape[-1]), "msda_zero_masked_rows_f32")
    return data


def cast_value_bf16(projected: torch.Tensor, mask) -> torch.Tensor:
    """bf16 copy of ``projected`` with the masked rows zeroed (no autograd)."""
    if not projected.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    _require(projected.dtype == torch.float32 and projected.is_contiguous(), "projected must be a contiguous fp32 tensor")
    m = None if mask is None else _mask_bytes(mask, projected.shape[:-1])
    out = torch.empty(projected.shape, dtype=torch.bfloat16, device=projected.device)
    rows = projected.numel() // projected.shape[-1] if projected.numel() else 0
    with _on_device(projected.device):
        _capi.check(_capi.lib.msda_value_prepare_bf16(_stream(projected.device), projected.data_ptr(),
                                                      None if m is None else m.data_ptr(), rows, projected.shape[-1],
                                                      out.data_ptr()), "msda_value_prepare_bf16")
    return out


class ValuePrepareFunction(Function):
    """``value.masked_fill(mask[..., None], 0)`` (+ optional bf16 cast) with its gradient."""

    @staticmethod
    def forward(ctx, projected, mask, to_bf16):
        ctx.has_mask = mask is not None
        if ctx.has_mask:
            ctx.save_for_backward(mask)
        if to_bf16:
            return cast_value_bf16(projected.contiguous(), mask)
        # fp32: the projection's output is not needed by its own backward (addmm saves its inputs), so the
        # masked rows are zeroed in place
        _require(ctx.has_mask, "nothing to do: no mask and no cast")
        _require(projected.is_contiguous(), "projected must be contiguous")
        ctx.mark_dirty(projected)
        return zero_masked_rows_(projected, mask)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_value):
        g = grad_value.float() if grad_value.dtype != torch.float32 else grad_value
        if ctx.has_mask:
            (mask,) = ctx.saved_tensors
            g = g.contiguous()
            if g.data_ptr() == grad_value.data_ptr():
                g = g.clone()  # the incoming gradient may be shared with other consumers
            zero_masked_rows_(g, mask)
        return g, None, None


def prepare_value(projected: torch.Tensor, mask=None, dtype=None) -> torch.Tensor:
    """The projected value ready for the sampling kernels: padded rows zeroed, optionally cast to bf16."""
    to_bf16 = dtype == torch.bfloat16 and projected.dtype == torch.float32
    if mask is None and not to_bf16:
        return projected if dtype is None else projected.to(dtype)
    if (projected.dtype != torch.float32 or dtype not in (None, torch.float32, torch.bfloat16)
            or (not to_bf16 and projected.is_leaf and projected.requires_grad)):  # no in-place on a leaf
        out = projected if mask is None else projected.masked_fill(mask[..., None], 0.0)
        return out if dtype is None else out.to(dtype)
    return ValuePrepareFunction.apply(projected, mask, to_bf16)


class EncoderProposalsFunction(Function):
    @staticmethod
    def forward(ctx, memory, mask, spatial_shapes, wh_base):
        if not memory.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        _require(memory.dim() == 3 and memory.dtype == torch.float32, "memory must be (N, S, C) fp32")
        memory = memory.contiguous()
        n, s, c = memory.shape
        m = None if mask is None else _mask_bytes(mask, (n, s))
        c_shapes = _shapes_host(spatial_shapes)
        levels = len(c_shapes) // 2
        opts = _capi.MsdaOpts()
        opts.struct_size = ctypes.sizeof(_capi.MsdaOpts)
        opts.spatial_shapes_host = ctypes.cast(c_shapes, _capi._i64p)
        out_mem = torch.empty_like(memory)
        out_prop = torch.empty(n, s, 4, dtype=torch.float32, device=memory.device)
        work = (torch.empty(n * levels * 2, dtype=torch.int32, device=memory.device)
                if m is not None and n * levels > 128 else None)  # small tables are recomputed in the kernel
        if wh_base is not None:
            wh_base = wh_base.detach().to(device=memory.device, dtype=torch.float32).contiguous()
            _require(wh_base.numel() == 2, "learnedwh must have 2 elements")
        with _on_device(memory.device):
            _capi.check(_capi.lib.msda_encoder_proposals_f32(
                _stream(memory.device), memory.data_ptr(), None if m is None else m.data_ptr(), None,
                None if wh_base is None else wh_base.data_ptr(), n, s, c, levels, out_mem.data_ptr(),
                out_prop.data_ptr(), None if work is None else work.data_ptr(), ctypes.byref(opts)),
                "msda_encoder_proposals_f32")
        ctx.save_for_backward(out_prop)
        ctx.mark_non_differentiable(out_prop)
        return out_mem, out_prop

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_mem, _grad_prop):
        (out_prop,) = ctx.saved_tensors
        n, s, _ = out_prop.shape
        grad_mem = grad_mem.contiguous()
        g = torch.empty_like(grad_mem)
        with _on_device(grad_mem.device):
            _capi.check(_capi.lib.msda_encoder_proposals_backward_f32(
                _stream(grad_mem.device), grad_mem.data_ptr(), out_prop.data_ptr(), n, s, grad_mem.shape[-1],
                g.data_ptr()), "msda_encoder_proposals_backward_f32")
        return g, None, None, None


def gen_encoder_output_proposals(memory, memory_padding_mask, spatial_shapes, learnedwh=None):
    """Drop-in for /root/reference/models/richsem/utils.py:10-65.

    memory (N, S, C) fp32, memory_padding_mask (N, S) bool, spatial_shapes (L, 2) -> (output_memory (N, S, C),
    output_proposals (N, S, 4)).  ``learnedwh`` (2,) is applied as in the reference but gets no gradient from
    this path (RichSem's config keeps two_stage_learn_wh off); a learnedwh that requires grad raises.
    """
    wh_base = None
    if learnedwh is not None:
        if learnedwh.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("gen_encoder_output_proposals: no gradient for learnedwh on the B200 path")
        wh_base = learnedwh.detach().sigmoid()
    return EncoderProposalsFunction.apply(memory, memory_padding_mask, spatial_shapes, wh_base)


def class_scores(enc_outputs_class: torch.Tensor) -> torch.Tensor:
    """``enc_outputs_class.max(-1)[0]`` (deformable_transformer.py:369) — (N, S, K) fp32 -> (N, S)."""
    if not enc_outputs_class.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    _require(enc_outputs_class.dtype == torch.float32 and enc_outputs_class.dim() >= 2, "class logits must be fp32 (..., K)")
    x = enc_outputs_class.detach().contiguous()
    out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    rows = out.numel()
    with _on_device(x.device):
        _capi.check(_capi.lib.msda_rowmax_f32(_stream(x.device), x.data_ptr(), rows, x.shape[-1], out.data_ptr()),
                    "msda_rowmax_f32")
    return out


def topk_rows(scores: torch.Tensor, k: int, return_values: bool = False):
    """``torch.topk(scores, k, dim=1)[1]`` for (N, S) fp32 scores: (N, k) int64, sorted by descending score (equal
    scores: lower index first).  k <= 1024."""
    if not scores.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    _require(scores.dtype == torch.float32 and scores.dim() == 2, "scores must be (N, S) fp32")
    _require(1 <= k <= scores.shape[1], f"selected index k out of range: k={k}, row length {scores.shape[1]}")
    s = scores.detach().contiguous()
    idx = torch.empty(s.shape[0], k, dtype=torch.int64, device=s.device)
    val = torch.empty(s.shape[0], k, dtype=torch.float32, device=s.device) if return_values else None
    with _on_device(s.device):
        _capi.check(_capi.lib.msda_topk_rows_f32(_stream(s.device), s.data_ptr(), s.shape[0], s.shape[1], k,
                                                 idx.data_ptr(), None if val is None else val.data_ptr()),
                    "msda_topk_rows_f32")
    return (val, idx) if return_values else idx


def topk_proposals(enc_outputs_class: torch.Tensor, topk: int) -> torch.Tensor:
    """deformable_transformer.py:367-369 in two kernels: ``torch.topk(enc_outputs_class.max(-1)[0], topk, dim=1)[1]``
    — the indices of the ``topk`` tokens with the largest best-class logit, (N, topk) int64.  Not differentiable
    (indices), like the reference expression."""
    return topk_rows(class_scores(enc_outputs_class), topk)


class AddLayerNormFunction(Function):
    """``F.layer_norm(x + residual, (C,), weight, bias, eps)`` in one kernel pass, with its gradients."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps):
        if not x.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        _require(x.dtype == torch.float32, "x must be fp32")
        _require(residual is None or (residual.shape == x.shape and residual.dtype == x.dtype and residual.is_cuda),
                 "residual must match x")
        c = x.shape[-1]
        _require(weight is None or (weight.numel() == c and weight.dtype == torch.float32), "weight must be (C,) fp32")
        _require(bias is None or (bias.numel() == c and bias.dtype == torch.float32), "bias must be (C,) fp32")
        x = x.contiguous()
        residual = None if residual is None else residual.contiguous()
        weight = None if weight is None else weight.contiguous()
        bias = None if bias is None else bias.contiguous()
        rows = x.numel() // c if x.numel() else 0
        out = torch.empty_like(x)
        stats = torch.empty(2, rows, dtype=torch.float32, device=x.device)  # mean, rstd
        with _on_device(x.device):
            _capi.check(_capi.lib.msda_add_layernorm_f32(
                _stream(x.device), x.data_ptr(), None if residual is None else residual.data_ptr(),
                None if weight is None else weight.data_ptr(), None if bias is None else bias.data_ptr(), rows, c,
                float(eps), out.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr()), "msda_add_layernorm_f32")
        ctx.save_for_backward(x, residual, weight, stats)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x, residual, weight, stats = ctx.saved_tensors
        c = x.shape[-1]
        rows = stats.shape[1]
        grad_out = grad_out.contiguous()
        grad_in = torch.empty_like(x)
        want_w = weight is not None and ctx.needs_input_grad[2]
        want_b = ctx.has_bias and ctx.needs_input_grad[3]
        gw = torch.empty(c, dtype=torch.float32, device=x.device) if want_w else None
        gb = torch.empty(c, dtype=torch.float32, device=x.device) if want_b else None
        ws_bytes = _capi.lib.msda_add_layernorm_workspace_bytes(rows, c) if (want_w or want_b) else 0
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
        with _on_device(x.device):
            _capi.check(_capi.lib.msda_add_layernorm_backward_f32(
                _stream(x.device), grad_out.data_ptr(), x.data_ptr(), None if residual is None else residual.data_ptr(),
                None if weight is None else weight.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), rows, c,
                grad_in.data_ptr(), None if gw is None else gw.data_ptr(), None if gb is None else gb.data_ptr(),
                None if ws is None else ws.data_ptr(), ws_bytes), "msda_add_layernorm_backward_f32")
        return (grad_in if ctx.needs_input_grad[0] else None,
                grad_in if residual is not None and ctx.needs_input_grad[1] else None, gw, gb, None)


def add_layer_norm_supported(x, residual=None) -> bool:
    """What the kernels serve: fp32 CUDA rows of 128 * k channels, k <= 4 (RichSem: d_model 256)."""
    c = x.shape[-1] if x.dim() else 0
    return (x.is_cuda and x.dtype == torch.float32 and c % 128 == 0 and 128 <= c <= 512
            and (residual is None or (residual.is_cuda and residual.dtype == torch.float32 and residual.shape == x.shape)))


def add_layer_norm(x, residual, norm: "torch.nn.LayerNorm | None" = None, weight=None, bias=None, eps=1e-5):
    """``norm(x + residual)`` (deformable_transformer.py:871-872): pass the ``nn.LayerNorm`` module, or weight / bias /
    eps.  Shapes the kernels do not serve take the two-kernel PyTorch expression."""
    if norm is not None:
        weight, bias, eps = norm.weight, norm.bias, norm.eps
    if not add_layer_norm_supported(x, residual):
        y = x if residual is None else x + residual
        return torch.nn.functional.layer_norm(y, (y.shape[-1],), weight, bias, eps)
    return AddLayerNormFunction.apply(x, residual, weight, bias, eps)
