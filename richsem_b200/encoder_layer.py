"""Deformable-encoder layer and 6-layer encoder stack around the B200 ``MSDeformAttn`` — the unit of BASELINE
config 4 ("batch-sharded encoder-layer train step, DDP grad all-reduce") and SURVEY section 8f row 3
(encoder-layer glue + CUDA-graph capture of the 6-layer encoder).

Mirrors the structure and parameter names of the reference
(/root/reference/models/richsem/deformable_transformer.py:825-881 ``DeformableTransformerEncoderLayer``: self_attn,
norm1, linear1, linear2, norm2; :470-618 ``TransformerEncoder``: ``layers``) so a reference state dict loads; dropout
and the optional channel attention are left out — the RichSem config trains with dropout 0.0
(config/RichSem/baseline_4scale.py:42).  The sampling core, its elementwise neighbours and the two residual +
LayerNorm epilogues (:871-872, :866-867; one pass each, csrc/msda_layernorm.cu) are this library's kernels; the Linear
layers are stock PyTorch (cuBLAS GEMMs).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .ops.functions.aux_functions import add_layer_norm
from .ops.modules import MSDeformAttn


class DeformableEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_points=4, value_dtype=None,
                 fuse_prologue=None, fuse_epilogue=True):
        super().__init__()
        self.fuse_epilogue = fuse_epilogue  # residual add + LayerNorm in one kernel (False: the PyTorch expressions)
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, value_dtype=value_dtype,
                                      fuse_prologue=fuse_prologue)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask=None):
        query = src if pos is None else src + pos
        src2 = self.self_attn(query, reference_points, src, spatial_shapes, level_start_index, key_padding_mask)
        if self.fuse_epilogue and src.is_cuda:
            src = add_layer_norm(src, src2, self.norm1)
            return add_layer_norm(src, self.linear2(F.relu(self.linear1(src))), self.norm2)
        src = self.norm1(src + src2)
        return self.norm2(src + self.linear2(F.relu(self.linear1(src))))


def encoder_reference_points(shapes, batch, device):
    """(N, S, L, 2): every token's pixel centre, replicated to all levels, valid_ratio = 1
    (deformable_transformer.py:512-525)."""
    from .synthetic import encoder_reference_points as pts

    ref = pts(shapes, device)  # (S, 2)
    return ref[None, :, None, :].expand(batch, ref.shape[0], len(shapes), 2).contiguous()


def get_reference_points(shapes, valid_ratios):
    """deformable_transformer.py:512-525 for host-known level shapes: token (y, x) of level l has the reference point
    ((x + 0.5) / (vr_x[l] * W_l), (y + 0.5) / (vr_y[l] * H_l)), expressed in every level k by multiplying with
    valid_ratios[:, k].  shapes: list of (H, W); valid_ratios (N, L, 2) as (x, y).  Returns (N, S, L, 2)."""
    dev = valid_ratios.device
    per_level = []
    for l, (h, w) in enumerate(shapes):
        ys = (torch.arange(h, dtype=torch.float32, device=dev) + 0.5).view(h, 1).expand(h, w).reshape(-1)
        xs = (torch.arange(w, dtype=torch.float32, device=dev) + 0.5).view(1, w).expand(h, w).reshape(-1)
        ry = ys[None] / (valid_ratios[:, None, l, 1] * h)
        rx = xs[None] / (valid_ratios[:, None, l, 0] * w)
        per_level.append(torch.stack((rx, ry), -1))
    ref = torch.cat(per_level, 1)
    return ref[:, :, None] * valid_ratios[:, None]


class DeformableEncoder(nn.Module):
    """The deformable encoder stack (deformable_transformer.py:470-618 with deformable_encoder=True,
    two_stage_type 'standard', no layer dropout, no final norm — RichSem's configuration): reference points are
    computed once and shared by all layers."""

    def __init__(self, num_layers=6, d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_points=4, value_dtype=None,
                 fuse_prologue=None, fuse_epilogue=True):
        super().__init__()
        self.layers = nn.ModuleList(DeformableEncoderLayer(d_model, d_ffn, n_levels, n_heads, n_points, value_dtype,
                                                           fuse_prologue, fuse_epilogue) for _ in range(num_layers))

    def forward(self, src, pos, spatial_shapes, level_start_index, valid_ratios, key_padding_mask=None,
                reference_points=None):
        if reference_points is None:
            from . import _capi

            meta = _capi.level_meta(spatial_shapes, level_start_index)  # host mirror, cached: no sync per call
            reference_points = get_reference_points(meta.shapes, valid_ratios)
        out = src
        for layer in self.layers:
            out = layer(out, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask)
        return out


class GraphedTrainStep:
    """One forward + backward of ``model`` captured in a CUDA graph and replayed.

    The MSDeformAttn kernels take 0.1-0.4 ms and the layers around them are dozens of short PyTorch kernels, so an
    eager encoder step is bounded by launch gaps; the library never synchronises or reads device memory on its hot
    path, so the whole step is capturable.  Inputs are copied into static buffers; parameter
    gradients land in the parameters' ``.grad`` (static as well), the loss in ``self.loss``.

        step = GraphedTrainStep(model, loss_fn, example_args)   # warms up, captures
        loss = step(*args)                                       # copy-in + replay
    """

    def __init__(self, model, loss_fn, example_args, warmup=3):
        self.model, self.loss_fn = model, loss_fn
        # Every tensor the kernels read on the device gets a static buffer that __call__ refreshes: floating-point
        # inputs and bool / uint8 masks.  Integer tensors (spatial_shapes, level_start_index) are different: the
        # library bakes their VALUES into the captured launches (kernel-parameter constant memory), so they are
        # captured by identity and __call__ refuses other values.
        self.static_args = [a.clone() if self._refreshed(a) else a for a in example_args]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._zero()
                self.loss_fn(self.model(*self.static_args)).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._zero()
        with torch.cuda.graph(self.graph):
            self.loss = self.loss_fn(self.model(*self.static_args))
            self.loss.backward()
        self.params = [p for p in self.model.parameters() if p.grad is not None]
        self.grads = [p.grad for p in self.params]  # static: every replay overwrites them

    def _zero(self):
        for p in self.model.parameters():
            p.grad = None

    @staticmethod
    def _refreshed(a):
        return isinstance(a, torch.Tensor) and (a.is_floating_point() or a.dtype in (torch.bool, torch.uint8))

    def __call__(self, *args):
        if len(args) != len(self.static_args):
            raise ValueError(f"expected {len(self.static_args)} arguments, got {len(args)}")
        for dst, src in zip(self.static_args, args):
            if self._refreshed(dst):
                if not isinstance(src, torch.Tensor) or src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError("argument does not match the captured tensor's shape / dtype")
                if src is not dst:
                    dst.copy_(src)
            elif isinstance(dst, torch.Tensor):
                if src is not dst and not (isinstance(src, torch.Tensor) and src.shape == dst.shape and torch.equal(src, dst)):
                    raise ValueError("integer tensor arguments (level shapes / start indices) are baked into the "
                                     "captured graph; capture a new GraphedTrainStep for different values")
            elif (src is None) != (dst is None):
                raise ValueError("an argument that was None at capture time must stay None (and vice versa)")
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g
        return self.loss
