"""Deformable-decoder layer around the B200 ``MSDeformAttn`` — the caller of BASELINE config 3 (decoder
cross-attention, 900 + 200 queries, 4-d reference boxes).

Mirrors the structure and parameter names of the reference's ``DeformableTransformerDecoderLayer``
(/root/reference/models/richsem/deformable_transformer.py:883-1068) in RichSem's configuration — ``decoder_sa_type='sa'``,
``module_seq=['sa', 'ca', 'ffn']``, relu, no key-aware projection, dropout 0.0 (config/RichSem/baseline_4scale.py:42) —
so a reference state dict loads: ``cross_attn``, ``norm1``, ``self_attn``, ``norm2``, ``linear1``, ``linear2``,
``norm3``.  Tensors are sequence-first as in the reference (``tgt``: (nq, bs, C), ``memory``: (S, bs, C)).

What runs on the library: the cross-attention's sampling core (split kernels; with ``fuse_prologue=True`` the softmax
and the box arithmetic of ms_deform_attn.py:98-111 too) and the three residual + LayerNorm epilogues (:1002-1003,
:973-974, :944-945).  Self-attention (``nn.MultiheadAttention``) and the Linear layers are stock PyTorch.
"""
from __future__ import annotations

import torch.nn.functional as F
from torch import nn

from .ops.functions.aux_functions import add_layer_norm
from .ops.modules import MSDeformAttn


class DeformableDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_points=4, value_dtype=None, fuse_prologue=None,
                 fuse_epilogue=True):
        super().__init__()
        self.fuse_epilogue = fuse_epilogue
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, value_dtype=value_dtype,
                                       fuse_prologue=fuse_prologue)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=0.0)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm3 = nn.LayerNorm(d_model)

    def _add_norm(self, x, y, norm):
        if self.fuse_epilogue and x.is_cuda:
            return add_layer_norm(x, y, norm)
        return norm(x + y)

    def forward(self, tgt, tgt_query_pos=None, tgt_reference_points=None, memory=None, memory_key_padding_mask=None,
                memory_level_start_index=None, memory_spatial_shapes=None, self_attn_mask=None):
        """tgt (nq, bs, C); tgt_query_pos (nq, bs, C) or None; tgt_reference_points (nq, bs, L, 2 | 4);
        memory (S, bs, C); memory_key_padding_mask (bs, S) bool or None -> (nq, bs, C)."""
        # self-attention (:965-974)
        q = k = tgt if tgt_query_pos is None else tgt + tgt_query_pos
        tgt = self._add_norm(tgt, self.self_attn(q, k, tgt, attn_mask=self_attn_mask)[0], self.norm2)
        # cross-attention (:998-1003)
        query = (tgt if tgt_query_pos is None else tgt + tgt_query_pos).transpose(0, 1)
        tgt2 = self.cross_attn(query, tgt_reference_points.transpose(0, 1).contiguous(), memory.transpose(0, 1),
                               memory_spatial_shapes, memory_level_start_index, memory_key_padding_mask).transpose(0, 1)
        tgt = self._add_norm(tgt, tgt2, self.norm1)
        # FFN (:941-945)
        return self._add_norm(tgt, self.linear2(F.relu(self.linear1(tgt))), self.norm3)
