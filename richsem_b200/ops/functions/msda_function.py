"""``MSDeformAttnFunction`` — autograd front end of the B200 MSDeformAttn kernels.

Keeps the call signature and gradient tuple of the reference's Function
(/root/reference/models/richsem/ops/functions/ms_deform_attn_func.py:21-38):

    MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
                               sampling_locations, attention_weights, im2col_step)

and returns gradients for value, sampling_locations and attention_weights only.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import MultiScaleDeformableAttention as _ext


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        result = _ext.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return result

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, starts, locations, weights = ctx.saved_tensors
        # the kernels index grad_output as (N, Lq, M, D) densely (reference asserts the same,
        # ms_deform_attn_cuda.cu:98)
        # (an input that needs no gradient — e.g. a detached / frozen `value` — skips its half of the work)
        g_value, g_loc, g_weight = _ext.ms_deform_attn_backward(
            value, shapes, starts, locations, weights, grad_output.contiguous(), ctx.im2col_step,
            _need_grad_value=ctx.needs_input_grad[0])
        if g_value is not None and g_value.dtype != value.dtype:  # bf16 value: autograd wants the input's dtype
            g_value = g_value.to(value.dtype)
        return g_value, None, None, g_loc, g_weight, None


class MSDeformAttnFusedFunction(Function):
    """MSDeformAttn with the module's prologue fused into the kernels.

        MSDeformAttnFusedFunction.apply(value, value_spatial_shapes, value_level_start_index,
                                        reference_points, sampling_offsets, attention_logits, im2col_step)

    takes the RAW outputs of the ``sampling_offsets`` / ``attention_weights`` Linear layers — (N, Lq, M, L, P, 2) and
    (N, Lq, M, L*P) — plus ``reference_points`` (N, Lq, L, 2 | 4), and computes the softmax and the sampling locations
    (/root/reference/models/richsem/ops/modules/ms_deform_attn.py:98-111) inside the kernels' decode step.  Gradients
    are returned for value, sampling_offsets and attention_logits; reference_points gets none (RichSem computes them
    without gradient: deformable_transformer.py:512-525 for the encoder, detached boxes in the decoder).
    Only available where ``MultiScaleDeformableAttention.fused_prologue_supported`` says so.
    """

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points, sampling_offsets,
                attention_logits, im2col_step):
        ctx.im2col_step = im2col_step
        result = _ext.ms_deform_attn_forward_fused(value, value_spatial_shapes, value_level_start_index, reference_points,
                                                   sampling_offsets, attention_logits, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, reference_points, sampling_offsets,
                              attention_logits)
        return result

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, starts, ref, offsets, logits = ctx.saved_tensors
        if ctx.needs_input_grad[3]:
            raise RuntimeError("MSDeformAttnFusedFunction does not differentiate reference_points; use the unfused path")
        g_value, g_off, g_logit = _ext.ms_deform_attn_backward_fused(
            value, shapes, starts, ref, offsets, logits, grad_output.contiguous(), ctx.im2col_step)
        if g_value.dtype != value.dtype:
            g_value = g_value.to(value.dtype)
        return g_value, None, None, None, g_off, g_logit, None
