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
