"""Synthetic MSDeformAttn workloads of the shapes RichSem produces.

Shapes follow the R50 backbone's stride arithmetic (reference: models/richsem/backbone.py:69-71,
155-156 and models/richsem/richsem.py:295-310); sampling-location distributions follow how the
reference builds them in the module (ops/modules/ms_deform_attn.py:64-70, 102-108) and the
transformer (deformable_transformer.py:512-525).  Used by bench.py, tests and smoke.
"""
from __future__ import annotations

import math

import torch

N_HEADS, HEAD_DIM, N_LEVELS, N_POINTS = 8, 32, 4, 4


def _half(x: int) -> int:  # 3x3 / stride 2 / pad 1 (and 7x7 s2 p3): floor((x-1)/2)+1
    return (x - 1) // 2 + 1


def level_shapes(height: int, width: int, n_levels: int = N_LEVELS):
    """(H_l, W_l) of the 4-scale R50 pyramid: strides 8, 16, 32 and the extra stride-64 conv."""
    h, w = height, width
    for _ in range(3):  # conv1, maxpool, layer2 -> stride 8
        h, w = _half(h), _half(w)
    shapes = [(h, w)]
    for _ in range(n_levels - 1):
        h, w = _half(h), _half(w)
        shapes.append((h, w))
    return shapes


def level_tensors(shapes, device):
    shp = torch.as_tensor(shapes, dtype=torch.long, device=device)
    hw = [h * w for h, w in shapes]
    starts = [0]
    for x in hw[:-1]:
        starts.append(starts[-1] + x)
    return shp, torch.as_tensor(starts, dtype=torch.long, device=device), sum(hw)


def encoder_reference_points(shapes, device, dtype=torch.float32):
    """Pixel centres of every token, normalised; valid_ratio 1 (deformable_transformer.py:512-525). (S, 2)"""
    pts = []
    for h, w in shapes:
        ys = (torch.arange(h, dtype=dtype, device=device) + 0.5) / h
        xs = (torch.arange(w, dtype=dtype, device=device) + 0.5) / w
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([gx.reshape(-1), gy.reshape(-1)], -1))
    return torch.cat(pts, 0)


def head_directions(n_heads=N_HEADS, device="cpu"):
    ang = torch.arange(n_heads, dtype=torch.float32, device=device) * (2.0 * math.pi / n_heads)
    d = torch.stack([ang.cos(), ang.sin()], -1)
    return d / d.abs().max(-1, keepdim=True)[0]


def locations_uniform(n, lq, gen, device, lo=0.0, hi=1.0, m=N_HEADS, l=N_LEVELS, p=N_POINTS):
    """Distribution U: uniform in [lo, hi)^2 (test.py:34 uses [0,1)); widen to exercise zero padding."""
    return torch.rand(n, lq, m, l, p, 2, generator=gen, device=device) * (hi - lo) + lo


def locations_encoder(n, shapes, gen, device, jitter_px=0.5, m=N_HEADS, p=N_POINTS):
    """Distribution E: token centres + the module's initial offsets (i+1)*dir_m pixels of each level
    + Gaussian jitter (keeps samples off the integer pixel lattice).  (N, S, M, L, P, 2)"""
    l = len(shapes)
    ref = encoder_reference_points(shapes, device)  # (S, 2)
    s = ref.shape[0]
    dirs = head_directions(m, device)  # (M, 2)
    steps = torch.arange(1, p + 1, dtype=torch.float32, device=device)
    off_px = dirs[:, None, None, :] * steps[None, None, :, None]  # (M,1,P,2)
    off_px = off_px.expand(m, l, p, 2)
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=device)  # (L,2)
    noise = torch.randn(n, s, m, l, p, 2, generator=gen, device=device) * jitter_px
    loc = ref[None, :, None, None, None, :] + (off_px[None, None] + noise) / wh[None, None, None, :, None, :]
    return loc.contiguous()


def locations_decoder(n, lq, gen, device, m=N_HEADS, l=N_LEVELS, p=N_POINTS):
    """Distribution Dn: reference boxes (centre uniform, w,h ~ U(0.02,0.6)), offsets ~ N(0, 2^2):
    loc = c + off / P * wh * 0.5 (ms_deform_attn.py:106-108).  A few % fall outside [0,1]."""
    c = torch.rand(n, lq, 1, 1, 1, 2, generator=gen, device=device)
    wh = torch.rand(n, lq, 1, 1, 1, 2, generator=gen, device=device) * 0.58 + 0.02
    off = torch.randn(n, lq, m, l, p, 2, generator=gen, device=device) * 2.0
    return (c + off / p * wh * 0.5).contiguous()


def attention_weights(n, lq, gen, device, m=N_HEADS, l=N_LEVELS, p=N_POINTS):
    logits = torch.randn(n, lq, m, l * p, generator=gen, device=device)
    return torch.softmax(logits, -1).view(n, lq, m, l, p).contiguous()


def make_inputs(kind, n, shapes, device, seed=1234, lq=None, dtype=torch.float32, m=N_HEADS, d=HEAD_DIM,
                p=N_POINTS):
    """kind: 'E' (encoder self-attention, Lq=S), 'U' (uniform), 'Dn' (decoder boxes).
    Returns dict(value, shapes, starts, loc, attw, grad_out, S, Lq)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    shp, starts, s = level_tensors(shapes, device)
    l = len(shapes)
    if kind == "E":
        lq = s
        loc = locations_encoder(n, shapes, gen, device, m=m, p=p)
    elif kind == "U":
        lq = lq or s
        loc = locations_uniform(n, lq, gen, device, m=m, l=l, p=p)
    elif kind == "Dn":
        lq = lq or 1100
        loc = locations_decoder(n, lq, gen, device, m=m, l=l, p=p)
    else:
        raise ValueError(kind)
    value = torch.randn(n, s, m, d, generator=gen, device=device)
    attw = attention_weights(n, lq, gen, device, m=m, l=l, p=p)
    grad_out = torch.randn(n, lq, m * d, generator=gen, device=device)
    return dict(value=value.to(dtype), shapes=shp, starts=starts, loc=loc, attw=attw,
                grad_out=grad_out.to(dtype), S=s, Lq=lq, shape_list=list(shapes))


def algorithmic_bytes(n, s, lq, m=N_HEADS, d=HEAD_DIM, l=N_LEVELS, p=N_POINTS, value_bytes=4, out_bytes=4):
    """Compulsory HBM bytes (each operand once), SURVEY.md §8(d).  Returns (fwd, bwd)."""
    val = n * s * m * d
    loc = n * lq * m * l * p * 2 * 4
    w = n * lq * m * l * p * 4
    out = n * lq * m * d
    fwd = val * value_bytes + loc + w + out * out_bytes
    bwd = out * out_bytes + val * value_bytes + loc + w + val * 4 + loc + w
    return fwd, bwd
