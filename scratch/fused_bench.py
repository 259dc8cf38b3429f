"""Fused prologue: time of the sampling core + its prologue, forward + backward, bs=2 encoder shape."""
import sys, torch
sys.path.insert(0, '.')
from richsem_b200 import synthetic as syn
from richsem_b200.ops.functions import MSDeformAttnFunction, MSDeformAttnFusedFunction
from richsem_b200.ops.modules import MSDeformAttn

dev = "cuda:0"
shapes = syn.level_shapes(800, 1333)
shp, starts, S = syn.level_tensors(shapes, dev)
g = torch.Generator(device=dev).manual_seed(3)
n, m, d, L, P = 2, 8, 32, 4, 4
value = torch.randn(n, S, m, d, generator=g, device=dev, requires_grad=True)
ref = syn.encoder_reference_points(shapes, dev)[None, :, None, :].expand(n, S, L, 2).contiguous()
offsets = (syn.head_directions(m, dev)[None, None, :, None, None, :] * torch.arange(1, P + 1, device=dev).view(1, 1, 1, 1, P, 1)
           + 0.5 * torch.randn(n, S, m, L, P, 2, generator=g, device=dev)).contiguous().requires_grad_(True)
logits = torch.randn(n, S, m, L * P, generator=g, device=dev, requires_grad=True)
grad_out = torch.randn(n, S, m * d, generator=g, device=dev)
norm = torch.stack([shp[..., 1], shp[..., 0]], -1)


def unfused():
    w = torch.softmax(logits, -1).view(n, S, m, L, P)
    loc = ref[:, :, None, :, None, :] + offsets / norm[None, None, None, :, None, :]
    MSDeformAttnFunction.apply(value, shp, starts, loc, w, 64).backward(grad_out)


def fused():
    MSDeformAttnFusedFunction.apply(value, shp, starts, ref, offsets, logits, 64).backward(grad_out)


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


tu, tf = timeit(unfused), timeit(fused)
print(f"sampling core + prologue, fwd+bwd, bs=2 encoder layer: unfused (PyTorch softmax / div / add around the op) {tu:.3f} ms, fused {tf:.3f} ms ({tu / tf:.2f}x)")
torch.manual_seed(0)
plain = MSDeformAttn(256, 4, 8, 4).to(dev)
fz = MSDeformAttn(256, 4, 8, 4, fuse_prologue=True).to(dev)
fz.load_state_dict(plain.state_dict())
src = torch.randn(n, S, 256, device=dev, requires_grad=True)
for name, mod in (("plain", plain), ("fuse_prologue", fz)):
    t = timeit(lambda: mod(src, ref, src, shp, starts, None).square().mean().backward())
    print(f"MSDeformAttn module fwd+bwd ({name}): {t:.3f} ms")

from richsem_b200 import MultiScaleDeformableAttention as ext
with torch.no_grad():
    w = torch.softmax(logits, -1).view(n, S, m, L, P).contiguous()
    loc = (ref[:, :, None, :, None, :] + offsets / norm[None, None, None, :, None, :]).contiguous()
    v, off, lg = value.detach(), offsets.detach(), logits.detach()
    print("kernels only: fwd unfused %.3f fused %.3f | bwd unfused %.3f fused %.3f ms" % (
        timeit(lambda: ext.ms_deform_attn_forward(v, shp, starts, loc, w, 64)),
        timeit(lambda: ext.ms_deform_attn_forward_fused(v, shp, starts, ref, off, lg, 64)),
        timeit(lambda: ext.ms_deform_attn_backward(v, shp, starts, loc, w, grad_out, 64)),
        timeit(lambda: ext.ms_deform_attn_backward_fused(v, shp, starts, ref, off, lg, grad_out, 64))))
    print("prologue alone (softmax + div + add): %.3f ms" % timeit(lambda: (torch.softmax(lg, -1), ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :])))
