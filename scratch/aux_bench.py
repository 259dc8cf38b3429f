"""Times the elementwise passes of csrc/msda_aux.cu against their PyTorch equivalents at the DINO shape
 and prints one JSON line per pass: algorithmic bytes, time, GB/s, fraction of the
measured HBM peak.  Inputs are rotated over enough distinct buffers to exceed L2.

    python scratch/aux_bench.py [--batch 2] [--iters 50]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from richsem_b200 import synthetic as syn  # noqa: E402
from richsem_b200.ops.functions import gen_encoder_output_proposals  # noqa: E402
from richsem_b200.ops.functions.aux_functions import cast_value_bf16, zero_masked_rows_  # noqa: E402


def peak():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def rect(shapes, frac):
    parts = []
    for h, w in shapes:
        m = torch.ones(h, w, dtype=torch.bool)
        m[: max(1, round(h * frac[0])), : max(1, round(w * frac[1]))] = False
        parts.append(m.reshape(-1))
    return torch.cat(parts)


def timeit(fns, iters):
    """fns: list of closures over distinct buffers (rotated).  Returns ms per call (CUDA events).  The calls are
    captured into one CUDA graph (one round over all buffers) and the graph is replayed, so that the figure is
    device time and not the Python launch rate (the passes take 10-30 us)."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            keep = [f() for f in fns]
    torch.cuda.current_stream().wait_stream(side)
    rounds = max(1, iters // len(fns))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / (rounds * len(fns))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--sets", type=int, default=8)
    ap.add_argument("--only", default="", help="'layernorm': just the residual + LayerNorm passes")
    a = ap.parse_args()
    dev = "cuda:0"
    shapes = syn.level_shapes(800, 1333)
    S = sum(h * w for h, w in shapes)
    n, c = a.batch, 256
    shp = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    mask = torch.stack([rect(shapes, (1.0, 1.0) if i % 2 == 0 else (0.8, 0.7)) for i in range(n)]).to(dev)
    masked_rows = int(mask.sum())
    xs = [torch.randn(n, S, c, device=dev) for _ in range(a.sets)]
    pk = peak()
    rows = n * S

    def report(name, ms, alg_bytes, torch_ms, note):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        print(json.dumps({"pass": name, "batch": n, "S": S, "C": c, "ms": ms, "algorithmic_bytes": alg_bytes,
                          "GBps": gbs, "frac_of_hbm_peak": gbs / pk, "peak_GBps": pk, "torch_ms": torch_ms,
                          "speedup_vs_torch": torch_ms / ms, "note": note}))

    # 8f-3 epilogue: norm(src + src2), forward (read 2 tensors, write 1) and backward (read 3, write 1)
    from richsem_b200.ops.functions.aux_functions import AddLayerNormFunction

    norm = torch.nn.LayerNorm(c).to(dev)
    rs = [torch.randn(n, S, c, device=dev) for _ in range(a.sets)]
    full = rows * c * 4
    ms = timeit([lambda x=x, r=r: AddLayerNormFunction.apply(x, r, norm.weight.detach(), norm.bias.detach(), 1e-5)
                 for x, r in zip(xs, rs)], a.iters)
    tms = timeit([lambda x=x, r=r: torch.nn.functional.layer_norm(x + r, (c,), norm.weight.detach(), norm.bias.detach())
                  for x, r in zip(xs, rs)], a.iters)
    report("add_layernorm_fwd", ms, 3 * full + rows * 8, tms, "vs F.layer_norm(x + r) (two kernels, five tensor passes)")

    def fwd_bwd(fn, x, r):
        x = x.detach().requires_grad_(True)
        r = r.detach().requires_grad_(True)
        norm.zero_grad(set_to_none=True)
        fn(x, r).backward(rs[0])
        return x.grad

    ms2 = timeit([lambda x=x, r=r: fwd_bwd(lambda u, v: AddLayerNormFunction.apply(u, v, norm.weight, norm.bias, 1e-5), x, r)
                  for x, r in zip(xs, rs)], a.iters)
    tms2 = timeit([lambda x=x, r=r: fwd_bwd(lambda u, v: norm(u + v), x, r) for x, r in zip(xs, rs)], a.iters)
    report("add_layernorm_fwd_bwd", ms2, 7 * full + 3 * rows * 8, tms2,
           "forward + backward incl. weight / bias gradients, vs norm(x + r) under autograd")
    # the backward pass alone, through the C ABI (no autograd bookkeeping around it)
    from richsem_b200 import _capi
    from richsem_b200.MultiScaleDeformableAttention import _stream

    stats = torch.empty(2, rows, device=dev)
    _capi.check(_capi.lib.msda_add_layernorm_f32(_stream(xs[0].device), xs[0].data_ptr(), rs[0].data_ptr(), norm.weight.data_ptr(),
                                                 norm.bias.data_ptr(), rows, c, 1e-5, torch.empty_like(xs[0]).data_ptr(),
                                                 stats[0].data_ptr(), stats[1].data_ptr()), "fwd")
    ws_bytes = _capi.lib.msda_add_layernorm_workspace_bytes(rows, c)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    gin = [torch.empty_like(x) for x in xs]
    gw, gb = torch.empty(c, device=dev), torch.empty(c, device=dev)

    def bwd_only(k):
        # statistics of set 0 for every set: timing only
        _capi.check(_capi.lib.msda_add_layernorm_backward_f32(
            _stream(xs[0].device), rs[(k + 1) % len(rs)].data_ptr(), xs[k].data_ptr(), rs[k].data_ptr(), norm.weight.data_ptr(),
            stats[0].data_ptr(), stats[1].data_ptr(), rows, c, gin[k].data_ptr(), gw.data_ptr(), gb.data_ptr(), ws.data_ptr(),
            ws_bytes), "bwd")

    ms3 = timeit([lambda k=k: bwd_only(k) for k in range(len(xs))], a.iters)
    report("add_layernorm_bwd_c_abi", ms3, 4 * full + 2 * rows * 8, float("nan"), "backward kernels alone (grad_in + weight / bias gradients)")
    del rs, gin
    if a.only == "layernorm":
        return
    # 8f-2, bf16: read fp32 (unmasked rows) + mask, write bf16
    alg = (rows - masked_rows) * c * 4 + rows + rows * c * 2
    ms = timeit([lambda x=x: cast_value_bf16(x, mask) for x in xs], a.iters)
    tms = timeit([lambda x=x: x.masked_fill(mask[..., None], 0.0).to(torch.bfloat16) for x in xs], a.iters)
    report("value_prepare_bf16", ms, alg, tms, "vs masked_fill + to(bf16)")
    # 8f-2, fp32 in place: read mask, write masked rows
    alg = rows + masked_rows * c * 4
    ms = timeit([lambda x=x: zero_masked_rows_(x, mask) for x in xs], a.iters)
    tms = timeit([lambda x=x: x.masked_fill(mask[..., None], 0.0) for x in xs], a.iters)
    report("zero_masked_rows_f32", ms, alg, tms, "vs out-of-place masked_fill")
    # 8f-4: read memory (kept rows) + mask, write memory + proposals
    keep = int(torch.isfinite(gen_encoder_output_proposals(xs[0], mask, shp)[1][..., 0]).sum())  # rows that are read
    alg = keep * c * 4 + rows + rows * c * 4 + rows * 16
    ms = timeit([lambda x=x: gen_encoder_output_proposals(x, mask, shp) for x in xs], a.iters)

    def torch_proposals(memory):
        # the reference's expression sequence (utils.py:10-65), restated for timing only
        props, cur = [], 0
        for lvl, (h, w) in enumerate(shapes):
            m = mask[:, cur:cur + h * w].view(n, h, w, 1)
            vh, vw = torch.sum(~m[:, :, 0, 0], 1), torch.sum(~m[:, 0, :, 0], 1)
            gy, gx = torch.meshgrid(torch.linspace(0, h - 1, h, dtype=torch.float32, device=dev),
                                    torch.linspace(0, w - 1, w, dtype=torch.float32, device=dev), indexing="ij")
            grid = torch.cat([gx.unsqueeze(-1), gy.unsqueeze(-1)], -1)
            scale = torch.cat([vw.unsqueeze(-1), vh.unsqueeze(-1)], 1).view(n, 1, 1, 2)
            grid = (grid.unsqueeze(0).expand(n, -1, -1, -1) + 0.5) / scale
            wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
            props.append(torch.cat((grid, wh), -1).view(n, -1, 4))
            cur += h * w
        p = torch.cat(props, 1)
        valid = ((p > 0.01) & (p < 0.99)).all(-1, keepdim=True)
        p = torch.log(p / (1 - p)).masked_fill(mask.unsqueeze(-1), float("inf")).masked_fill(~valid, float("inf"))
        om = memory.masked_fill(mask.unsqueeze(-1), 0.0).masked_fill(~valid, 0.0)
        return om, p

    tms = timeit([lambda x=x: torch_proposals(x) for x in xs], a.iters)
    report("encoder_proposals_f32", ms, alg, tms, "vs the reference's PyTorch expression sequence (utils.py:10-65)")

    # 8f-4, second half: best-class logit per token + top-900 tokens per image (deformable_transformer.py:367-369)
    from richsem_b200.ops.functions import topk_proposals
    from richsem_b200.ops.functions.aux_functions import class_scores, topk_rows

    for kc in (91, 1203):
        ls = [torch.randn(n, S, kc, device=dev) for _ in range(max(2, min(a.sets, 4)))]
        alg = rows * kc * 4 + rows * 4
        ms = timeit([lambda x=x: class_scores(x) for x in ls], a.iters)
        tms = timeit([lambda x=x: x.max(-1)[0] for x in ls], a.iters)
        report(f"class_scores_K{kc}", ms, alg, tms, "vs logits.max(-1)[0]")
        ms = timeit([lambda x=x: topk_proposals(x, 900) for x in ls], a.iters)
        tms = timeit([lambda x=x: torch.topk(x.max(-1)[0], 900, dim=1)[1] for x in ls], a.iters)
        report(f"topk_proposals_K{kc}", ms, alg + rows * 4 + n * 900 * 8, tms,
               "vs torch.topk(logits.max(-1)[0], 900, dim=1)[1]")
        del ls
    sc = [torch.randn(n, S, device=dev) for _ in range(4)]
    ms = timeit([lambda x=x: topk_rows(x, 900) for x in sc], a.iters)
    tms = timeit([lambda x=x: torch.topk(x, 900, dim=1)[1] for x in sc], a.iters)
    report("topk_rows_k900", ms, rows * 4 + n * 900 * 8, tms, "vs torch.topk(scores, 900, dim=1)[1] (latency-bound: one block per image)")


if __name__ == "__main__":
    main()
