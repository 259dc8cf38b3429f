"""Deformable-encoder layer around the B200 ``MSDeformAttn`` — the unit of BASELINE config 4
("batch-sharded encoder-layer train step, DDP grad all-reduce").

Mirrors the structure and parameter names of the reference layer
(/root/reference/models/richsem/deformable_transformer.py:825-881: self_attn, norm1, linear1,
linear2, norm2; dropout and the optional channel attention are left out — the RichSem config
trains with dropout 0.0, config/RichSem/baseline_4scale.py:42) so a reference state dict loads.
Everything but the sampling core is stock PyTorch (cuBLAS GEMMs, LayerNorm).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .ops.modules import MSDeformAttn


class DeformableEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_points=4, value_dtype=None):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, value_dtype=value_dtype)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask=None):
        query = src if pos is None else src + pos
        src = self.norm1(src + self.self_attn(query, reference_points, src, spatial_shapes, level_start_index,
                                              key_padding_mask))
        return self.norm2(src + self.linear2(F.relu(self.linear1(src))))


def encoder_reference_points(shapes, batch, device):
    """(N, S, L, 2): every token's pixel centre, replicated to all levels, valid_ratio = 1
    (deformable_transformer.py:512-525)."""
    from .synthetic import encoder_reference_points as pts

    ref = pts(shapes, device)  # (S, 2)
    return ref[None, :, None, :].expand(batch, ref.shape[0], len(shapes), 2).contiguous()
