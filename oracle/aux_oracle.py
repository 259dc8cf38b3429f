"""CPU oracle for the elementwise passes either side of the sampling core.

TEST INFRASTRUCTURE ONLY — imported by tests/ (and nothing in richsem_b200/).  Own-words restatements:

* ``value_prepare``      /root/reference/models/richsem/ops/modules/ms_deform_attn.py:94-97
* ``encoder_proposals``  /root/reference/models/richsem/utils.py:10-65 (gen_encoder_output_proposals)
* ``add_layer_norm``     /root/reference/models/richsem/deformable_transformer.py:871-872, 866-867
                         (``src = src + dropout(src2); src = norm(src)``, dropout 0)

Pinned: ``encoder_proposals`` is bit-identical to the reference function loaded by path in the build
container (tests/test_oracle.py::test_aux_oracle_matches_the_reference_function) and to the golden vectors in
tests/golden/proposals_*.npz, which tests/golden/make_golden_aux.py generated from the reference function.
"""
from __future__ import annotations

import importlib.util
from pathlib import Path

import torch


def value_prepare(projected, mask=None, dtype=None):
    """ms_deform_attn.py:94-97: padded tokens of the projected value are zero; optional storage cast
    (round-to-nearest-even, what ``Tensor.to(torch.bfloat16)`` does)."""
    out = projected.clone()
    if mask is not None:
        out[mask.bool()] = 0
    return out if dtype is None else out.to(dtype)


def value_prepare_backward(grad_value, mask=None):
    g = grad_value.float().clone()
    if mask is not None:
        g[mask.bool()] = 0
    return g


def encoder_proposals(memory, mask, spatial_shapes, learnedwh=None):
    """utils.py:10-65 restated token by token instead of level-wise meshgrids.

    For token (row y, column x) of level l in image b:
        cx = (x + 0.5) / valid_W[b,l]      cy = (y + 0.5) / valid_H[b,l]          (utils.py:27-38)
        w = base_w * 2**l                  h = base_h * 2**l                      (utils.py:39-43)
    with valid_H / valid_W = number of unpadded tokens in the level's first column / row, base = sigmoid(learnedwh)
    or 0.05.  A proposal is valid when all four numbers are in (0.01, 0.99) (utils.py:49); output proposals are
    logits log(p / (1 - p)), +inf where padded or invalid (utils.py:50-52); the memory row is zero there (:54-56).
    """
    n, s, _ = memory.shape
    shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if isinstance(spatial_shapes, torch.Tensor) else spatial_shapes)]
    if mask is None:
        mask = torch.zeros(n, s, dtype=torch.bool)
    mask = mask.bool()
    if learnedwh is not None:
        base = learnedwh.detach().float().sigmoid()
    else:
        base = torch.tensor([0.05, 0.05], dtype=torch.float32)
    prop = torch.empty(n, s, 4, dtype=torch.float32)
    cur = 0
    for l, (h, w) in enumerate(shapes):
        m = mask[:, cur:cur + h * w].view(n, h, w)
        valid_h = (~m[:, :, 0]).sum(1).to(torch.float32)          # (n,)
        valid_w = (~m[:, 0, :]).sum(1).to(torch.float32)
        xs = torch.arange(w, dtype=torch.float32) + 0.5
        ys = torch.arange(h, dtype=torch.float32) + 0.5
        cx = (xs[None, None, :] / valid_w[:, None, None]).expand(n, h, w)
        cy = (ys[None, :, None] / valid_h[:, None, None]).expand(n, h, w)
        bw = (base[0] * torch.tensor(2.0 ** l, dtype=torch.float32)).expand(n, h, w)
        bh = (base[1] * torch.tensor(2.0 ** l, dtype=torch.float32)).expand(n, h, w)
        prop[:, cur:cur + h * w] = torch.stack([cx, cy, bw, bh], -1).reshape(n, h * w, 4)
        cur += h * w
    assert cur == s
    valid = ((prop > 0.01) & (prop < 0.99)).all(-1)
    keep = valid & ~mask
    logit = torch.log(prop / (1 - prop))
    out_prop = torch.where(keep[..., None], logit, torch.full_like(logit, float("inf")))
    out_mem = torch.where(keep[..., None], memory, torch.zeros_like(memory))
    return out_mem, out_prop


def load_reference_proposals(reference_root="/root/reference"):
    """The reference's own gen_encoder_output_proposals, loaded by path (build container only)."""
    path = Path(reference_root) / "models" / "richsem" / "utils.py"
    if not path.exists():
        return None
    spec = importlib.util.spec_from_file_location("_richsem_ref_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.gen_encoder_output_proposals


def add_layer_norm(x, residual, weight, bias, eps=1e-5, dtype=torch.float64):
    """deformable_transformer.py:871-872: ``norm(src + src2)`` with nn.LayerNorm's definition (biased variance,
    ``(y - mean) / sqrt(var + eps) * weight + bias``), written out in ``dtype`` (fp64 by default: the kernels are
    compared against the exact result, torch's own fp32 kernel is a second witness in the tests)."""
    y = x.to(dtype) if residual is None else x.to(dtype) + residual.to(dtype)
    mean = y.mean(-1, keepdim=True)
    var = ((y - mean) ** 2).mean(-1, keepdim=True)
    out = (y - mean) / torch.sqrt(var + eps)
    if weight is not None:
        out = out * weight.to(dtype)
    if bias is not None:
        out = out + bias.to(dtype)
    return out
