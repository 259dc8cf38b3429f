"""Generates tests/golden/*.npz from the REFERENCE's own ms_deform_attn_core_pytorch
(/root/reference/models/richsem/ops/functions/ms_deform_attn_func.py:41-61), loaded by path in the
build container (the reference cannot travel to the GPU box).  Forward outputs come from the reference
function; gradients from torch.autograd through the same reference function.

    python tests/golden/make_golden.py          # rewrites the fixtures (CPU, seconds)

Cases
  tiny_f32 / tiny_f64   the reference's only test case (ops/test.py:21-36, seed 3, shapes (6,4),(3,2))
  enc_small_f32         D=32, M=8, L=4, P=4 encoder-style (Lq = S), distribution E, two images
  pad_small_f32         same shape, uniform locations widened to [-0.25, 1.25): zero padding, skipped samples
  dec_small_f32         decoder-style boxes (distribution Dn), Lq=37
  odd_dims_f64          D=5, M=3, L=3, P=2 (generic-kernel territory), fp64
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.msda_oracle import load_reference_core  # noqa: E402

HERE = Path(__file__).resolve().parent


def fwd_bwd(ref, value, shapes, loc, attw, grad_out):
    v = value.clone().requires_grad_(True)
    l = loc.clone().requires_grad_(True)
    a = attw.clone().requires_grad_(True)
    out = ref(v, torch.as_tensor(shapes, dtype=torch.long), l, a)
    out.backward(grad_out)
    return out.detach(), v.grad, l.grad, a.grad


def save(name, shapes, value, loc, attw, grad_out, out, gv, gl, ga):
    np.savez_compressed(HERE / f"{name}.npz", shapes=np.asarray(shapes, dtype=np.int64), value=value.numpy(),
                        loc=loc.numpy(), attw=attw.numpy(), grad_out=grad_out.numpy(), out=out.numpy(),
                        grad_value=gv.numpy(), grad_loc=gl.numpy(), grad_attw=ga.numpy())
    print(f"{name}: out {tuple(out.shape)} |out|max {out.abs().max():.3e}")


def jitter_off_lattice(loc, shapes, eps=2e-3):
    """Move samples whose pixel coordinate is within eps of an integer (bilinear kink, where
    d out / d loc is discontinuous and fp32 rounding picks the side) away from it."""
    loc = loc.clone()
    for l, (h, w) in enumerate(shapes):
        for ax, size in ((0, w), (1, h)):
            pix = loc[:, :, :, l, :, ax] * size - 0.5
            frac = pix - torch.floor(pix)
            near = (frac < eps) | (frac > 1 - eps)
            loc[:, :, :, l, :, ax] = torch.where(near, loc[:, :, :, l, :, ax] + 3 * eps / size, loc[:, :, :, l, :, ax])
    return loc


def main():
    ref = load_reference_core()
    assert ref is not None, "/root/reference is required to regenerate the golden vectors"
    from richsem_b200 import synthetic as syn

    # --- the reference's own test case --------------------------------------------------
    shapes = [(6, 4), (3, 2)]
    n, m, d, lq, nl, p = 1, 2, 2, 2, 2, 2
    s = sum(h * w for h, w in shapes)
    torch.manual_seed(3)
    value = torch.rand(n, s, m, d) * 0.01
    loc = torch.rand(n, lq, m, nl, p, 2)
    attw = torch.rand(n, lq, m, nl, p) + 1e-5
    attw /= attw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    grad_out = torch.randn(n, lq, m * d)
    save("tiny_f32", shapes, value, loc, attw, grad_out, *fwd_bwd(ref, value, shapes, loc, attw, grad_out))
    save("tiny_f64", shapes, value.double(), loc.double(), attw.double(), grad_out.double(),
         *fwd_bwd(ref, value.double(), shapes, loc.double(), attw.double(), grad_out.double()))

    # --- D=32 cases ------------------------------------------------------------------------
    shapes = [(8, 11), (4, 6), (2, 3), (1, 2)]
    i = syn.make_inputs("E", 2, shapes, "cpu", seed=11)
    loc = jitter_off_lattice(i["loc"], shapes)
    save("enc_small_f32", shapes, i["value"], loc, i["attw"], i["grad_out"],
         *fwd_bwd(ref, i["value"], shapes, loc, i["attw"], i["grad_out"]))

    gen = torch.Generator().manual_seed(12)
    i = syn.make_inputs("U", 2, shapes, "cpu", seed=12, lq=41)
    loc = jitter_off_lattice(syn.locations_uniform(2, 41, gen, "cpu", lo=-0.25, hi=1.25), shapes)
    save("pad_small_f32", shapes, i["value"], loc, i["attw"], i["grad_out"],
         *fwd_bwd(ref, i["value"], shapes, loc, i["attw"], i["grad_out"]))

    i = syn.make_inputs("Dn", 2, shapes, "cpu", seed=13, lq=37)
    loc = jitter_off_lattice(i["loc"], shapes)
    save("dec_small_f32", shapes, i["value"], loc, i["attw"], i["grad_out"],
         *fwd_bwd(ref, i["value"], shapes, loc, i["attw"], i["grad_out"]))

    # --- odd dims, fp64 -------------------------------------------------------------------
    shapes = [(5, 7), (3, 4), (2, 2)]
    gen = torch.Generator().manual_seed(14)
    n, m, d, lq, nl, p = 2, 3, 5, 9, 3, 2
    s = sum(h * w for h, w in shapes)
    value = torch.randn(n, s, m, d, generator=gen, dtype=torch.float64)
    loc = torch.rand(n, lq, m, nl, p, 2, generator=gen, dtype=torch.float64) * 1.3 - 0.15
    attw = torch.softmax(torch.randn(n, lq, m, nl * p, generator=gen, dtype=torch.float64), -1).view(n, lq, m, nl, p)
    grad_out = torch.randn(n, lq, m * d, generator=gen, dtype=torch.float64)
    save("odd_dims_f64", shapes, value, loc, attw, grad_out, *fwd_bwd(ref, value, shapes, loc, attw, grad_out))


if __name__ == "__main__":
    main()
