"""``MSDeformAttn`` — the multi-scale deformable attention layer used by RichSem's deformable
encoder (self-attention over the flattened feature pyramid) and decoder (cross-attention).

Drop-in for /root/reference/models/richsem/ops/modules/ms_deform_attn.py:30-115: same constructor
``(d_model=256, n_levels=4, n_heads=8, n_points=4)``, same forward signature, same parameter names
(``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj``) so published RichSem
checkpoints load unchanged, same initialisation.  The four Linear layers and the softmax stay in
PyTorch (north_star); only the sampling core runs on the hand-written kernels.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

import os

from ... import MultiScaleDeformableAttention as _ext
from ... import _capi
from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction, prepare_value


def _power_of_two(n) -> bool:
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


class MSDeformAttn(nn.Module):
    """Multi-scale deformable attention.

    Args:
        d_model:  hidden size C
        n_levels: feature levels L
        n_heads:  attention heads M (C must be divisible by M)
        n_points: sampling points P per head per level
    Extra (not in the reference):
        value_dtype: ``None`` keeps the reference behaviour (value in the input dtype);
            ``torch.bfloat16`` stores the projected value and the sampled output in bf16
            (fp32 accumulation; the kernels return fp32 gradients, and autograd hands value_proj a grad_value rounded to
            this dtype).
        fuse_prologue: compute the softmax and the sampling locations inside the kernels
            (``MSDeformAttnFusedFunction``) wherever the library supports it — encoder self-attention and
            decoder-sized cross-attention calls — instead of five elementwise PyTorch kernels around the op.  Same results within
            the op's tolerances; ``None`` reads ``MSDA_B200_FUSE_PROLOGUE`` (default off).
    """

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, value_dtype=None, fuse_prologue=None):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        if not _power_of_two(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention "
                          "head a power of 2 which is more efficient in our CUDA implementation.")
        self.im2col_step = 64  # kept for interface parity; the B200 kernels take the whole batch in one launch
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.value_dtype = value_dtype
        self.fuse_prologue = (os.environ.get("MSDA_B200_FUSE_PROLOGUE", "0") not in ("", "0")
                              if fuse_prologue is None else bool(fuse_prologue))

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # ms_deform_attn.py:62-76 — zero offset weights; offset bias = the head's compass direction
        # (unit in the max-norm) scaled by 1..P; uniform attention; xavier projections.
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
            direction = torch.stack([angle.cos(), angle.sin()], dim=-1)
            direction = direction / direction.abs().max(dim=-1, keepdim=True)[0]
            bias = direction.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
            bias = bias * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
            self.sampling_offsets.bias.copy_(bias.reshape(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """
        query                    (N, Lq, C)
        reference_points         (N, Lq, L, 2) in [0,1] (top-left (0,0), bottom-right (1,1), padding included)
                                 or (N, Lq, L, 4): (cx, cy, w, h) reference boxes
        input_flatten            (N, S, C) with S = sum_l H_l*W_l
        input_spatial_shapes     (L, 2) rows (H_l, W_l)
        input_level_start_index  (L,)
        input_padding_mask       (N, S) bool, True on padding; or None
        returns                  (N, Lq, C)
        """
        n, len_q, _ = query.shape
        n, len_in, _ = input_flatten.shape
        # ms_deform_attn.py:92 reads the device tensor here (a device->host sync per call, and illegal under CUDA-graph
        # capture); the same check runs on the cached host mirror of the level table
        if input_spatial_shapes.is_cuda:
            assert _capi.level_meta(input_spatial_shapes, input_level_start_index).spatial_size_sum == len_in
        else:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == len_in

        heads, levels, points = self.n_heads, self.n_levels, self.n_points
        # ms_deform_attn.py:94-97: projection, padded tokens zeroed, (N, S, M, D) view — the zeroing (in place for
        # fp32) and the optional bf16 cast are one pass of csrc/msda_aux.cu
        value = prepare_value(self.value_proj(input_flatten), input_padding_mask, self.value_dtype)
        value = value.view(n, len_in, heads, self.d_model // heads)
        offsets = self.sampling_offsets(query).view(n, len_q, heads, levels, points, 2)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")
        fused_value = value
        if (self.fuse_prologue and not reference_points.requires_grad and query.dtype == torch.float32
                and _ext.fused_prologue_supported(fused_value, levels, len_q, points)):
            logits = self.attention_weights(query).view(n, len_q, heads, levels * points)
            sampled = MSDeformAttnFusedFunction.apply(fused_value, input_spatial_shapes, input_level_start_index,
                                                      reference_points.contiguous(), offsets, logits, self.im2col_step)
            return self.output_proj(sampled.to(query.dtype))
        weights = F.softmax(self.attention_weights(query).view(n, len_q, heads, levels * points), -1)
        weights = weights.view(n, len_q, heads, levels, points)
        if reference_points.shape[-1] == 2:
            # offsets are in pixels of each level: normalise by (W_l, H_l)
            wh = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            locations = reference_points[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            # offsets are fractions of half the reference box size
            locations = (reference_points[:, :, None, :, None, :2]
                         + offsets / points * reference_points[:, :, None, :, None, 2:] * 0.5)
        else:
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")

        if self.value_dtype is not None and value.dtype != query.dtype:
            sampled = MSDeformAttnFunction.apply(value, input_spatial_shapes,
                                                 input_level_start_index, locations.float(), weights.float(),
                                                 self.im2col_step).to(query.dtype)
        else:
            sampled = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, locations,
                                                 weights, self.im2col_step)
        return self.output_proj(sampled)
