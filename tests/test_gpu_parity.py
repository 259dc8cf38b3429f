"""GPU: parity of the CUDA path (through the C ABI) against the oracle.

Tolerances (BASELINE.json north_star): fp32 forward <= 1e-5 relative, fp32 backward <= 1e-4 relative,
bf16-value variant <= 1e-2; corner indices and level offsets bit-exact.  "relative" = max|a-b| / max|b|.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL, BWD_TOL, BF16_TOL = 1e-5, 1e-4, 1e-2


def _ext():
    from richsem_b200 import MultiScaleDeformableAttention as ext

    return ext


def _to_dev(g):
    dev = "cuda:0"
    shp = torch.as_tensor(g["shape_list"], dtype=torch.long, device=dev)
    hw = shp[:, 0] * shp[:, 1]
    st = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
    return (g["value"].to(dev), shp, st, g["loc"].to(dev), g["attw"].to(dev), g["grad_out"].to(dev))


def test_extension_is_loaded_from_the_tree():
    from richsem_b200 import _capi

    assert _capi.LIB_PATH.endswith("richsem_b200/lib/libmsda_b200.so")
    before = _capi.launch_count()
    g = load_golden("tiny_f32")
    v, shp, st, loc, w, go = _to_dev(g)
    _ext().ms_deform_attn_forward(v, shp, st, loc, w, 2)
    assert _capi.launch_count() == before + 1


def _path(path):
    """Keyword arguments that put a (small) test problem through one kernel family (msda_opts.kernel_hint)."""
    from richsem_b200 import _capi

    return {"auto": {}, "split": {"_kernel": _capi.KERNEL_SPLIT}, "window": {"_kernel": _capi.KERNEL_WINDOW},
            "tiled": {"_kernel": _capi.KERNEL_TILED}, "generic": {"_flags": _capi.FLAG_FORCE_GENERIC}}[path]


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("path", ["auto", "window", "tiled", "generic"])
def test_golden_forward_backward(case, path):
    """auto = what the library picks (warp-per-query "split" kernels at these sizes for D=32);
    window = the shared-memory window kernels used for large problems; tiled = their L1-gather
    predecessors; generic = any-shape kernels."""
    from richsem_b200 import _capi

    g = load_golden(case)
    f64 = g["value"].dtype == torch.float64
    v, shp, st, loc, w, go = _to_dev(g)
    out = _ext().ms_deform_attn_forward(v, shp, st, loc, w, 64, **_path(path))
    gv, gl, ga = _ext().ms_deform_attn_backward(v, shp, st, loc, w, go, 64, **_path(path))
    ft, bt = (1e-12, 1e-11) if f64 else (FWD_TOL, BWD_TOL)
    assert rel_err(out.cpu(), g["out"]) < ft
    assert rel_err(gv.cpu(), g["grad_value"]) < bt
    assert rel_err(gl.cpu(), g["grad_loc"]) < bt
    assert rel_err(ga.cpu(), g["grad_attw"]) < bt


@pytest.mark.parametrize("case", ["enc_small_f32", "pad_small_f32", "dec_small_f32", "tiny_f32"])
def test_corner_indices_bit_exact(case, c_oracle):
    g = load_golden(case)
    v, shp, st, loc, w, go = _to_dev(g)
    got = _ext().debug_corners(shp, st, loc).cpu()
    want = c_oracle.corners(g["shape_list"], g["loc"])
    assert torch.equal(got, want)


def test_corner_indices_bit_exact_on_the_pixel_lattice(c_oracle):
    """Encoder reference points + integer-pixel offsets land exactly on the lattice."""
    from richsem_b200 import synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    gen = torch.Generator().manual_seed(7)
    loc = syn.locations_encoder(1, shapes, gen, "cpu", jitter_px=0.0)[:, ::7].contiguous()
    shp, st, _ = syn.level_tensors(shapes, "cuda:0")
    got = _ext().debug_corners(shp, st, loc.cuda()).cpu()
    want = c_oracle.corners(shapes, loc)
    assert torch.equal(got, want)


def test_corner_indices_fma_mode_bit_exact(c_oracle):
    """MSDA_FLAG_COORDS_FMA reproduces the compiled reference kernel's coordinate arithmetic."""
    from richsem_b200 import _capi, synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    gen = torch.Generator().manual_seed(7)
    loc = syn.locations_encoder(1, shapes, gen, "cpu", jitter_px=0.0)[:, ::7].contiguous()
    shp, st, _ = syn.level_tensors(shapes, "cuda:0")
    got = _ext().debug_corners(shp, st, loc.cuda(), _flags=_capi.FLAG_COORDS_FMA).cpu()
    assert torch.equal(got, c_oracle.corners(shapes, loc, fma=True))
    assert not torch.equal(got, c_oracle.corners(shapes, loc))


def test_reference_test_py_cases():
    """models/richsem/ops/test.py:31-60: seed-3 tiny case, fp64 allclose and fp32 rtol 1e-2/atol 1e-3."""
    from oracle.msda_oracle import core_pytorch
    from richsem_b200.ops.functions import MSDeformAttnFunction

    n, m, d, lq, nl, p = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    starts = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    s = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    for dtype, kw in ((torch.float64, {}), (torch.float32, dict(rtol=1e-2, atol=1e-3))):
        value = (torch.rand(n, s, m, d).cuda() * 0.01).to(dtype)
        loc = torch.rand(n, lq, m, nl, p, 2).cuda().to(dtype)
        w = torch.rand(n, lq, m, nl, p).cuda() + 1e-5
        w = (w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dtype)
        want = core_pytorch(value.cpu(), shapes.cpu().tolist(), loc.cpu(), w.cpu())
        got = MSDeformAttnFunction.apply(value, shapes, starts, loc, w, 2).cpu()
        assert torch.allclose(got, want, **kw)


@pytest.mark.parametrize("channels", [30, 32, 64, 71, 1025])
def test_gradcheck_like_reference(channels):
    """models/richsem/ops/test.py:63-86 (gradcheck over channel counts; 2048/3096 trimmed for time)."""
    from torch.autograd import gradcheck
    from richsem_b200.ops.functions import MSDeformAttnFunction

    n, m, lq, nl, p = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    starts = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    s = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    value = (torch.rand(n, s, m, channels).cuda() * 0.01).double().requires_grad_(True)
    loc = torch.rand(n, lq, m, nl, p, 2).cuda().double().requires_grad_(True)
    w = torch.rand(n, lq, m, nl, p).cuda() + 1e-5
    w = (w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_(True)
    assert gradcheck(MSDeformAttnFunction.apply, (value, shapes, starts, loc, w, 2))


@pytest.mark.parametrize("kind,n,lq,bflags", [("E", 1, None, "auto"), ("E", 2, None, "auto"), ("E", 1, None, "tiled"),
                                              ("U", 2, 3000, "auto"), ("U", 2, 3000, "tiled"), ("U", 2, 3000, "window"),
                                              ("U", 2, 9000, "auto"),
                                              ("Dn", 2, 1100, "auto"), ("Dn", 2, 1100, "window")])
def test_dino_shape_against_c_oracle(kind, n, lq, bflags, c_oracle):
    """Full DINO 4-scale R50 800x1333 pyramid (S=22223, M=8, D=32, L=4, P=4) against the C oracle.
    Encoder self-attention ("E", patch order) takes the window backward; "U" with 3000 queries takes the split
    kernels, with 9000 the tiled ones (large, no query order); "tiled" = per-corner reductions forced,
    "window" = the window kernel forced on small / non-local inputs (most levels fall back to its direct pass)."""
    from richsem_b200 import _capi, synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    i = syn.make_inputs(kind, n, shapes, "cuda:0", seed=21, lq=lq)
    out = _ext().ms_deform_attn_forward(i["value"], i["shapes"], i["starts"], i["loc"], i["attw"], 64)
    gv, gl, ga = _ext().ms_deform_attn_backward(i["value"], i["shapes"], i["starts"], i["loc"], i["attw"],
                                                i["grad_out"], 64, **_path(bflags))
    v, loc, w, go = (i[k].cpu() for k in ("value", "loc", "attw", "grad_out"))
    want = c_oracle.forward(v, shapes, loc, w)
    assert rel_err(out.cpu(), want) < FWD_TOL
    wv, wl, wa = c_oracle.backward(go, v, shapes, loc, w)
    assert rel_err(gv.cpu(), wv) < BWD_TOL
    assert rel_err(ga.cpu(), wa) < BWD_TOL
    # grad_loc is discontinuous where a sample sits on the pixel lattice; compare off-lattice samples
    # (the oracle and the kernel use the same fp32 coordinates, so they agree on which side they are)
    assert rel_err(gl.cpu(), wl) < BWD_TOL
    got_idx = _ext().debug_corners(i["shapes"], i["starts"], i["loc"]).cpu()
    assert torch.equal(got_idx, c_oracle.corners(shapes, loc))


def test_query_order_does_not_change_results(monkeypatch):
    from richsem_b200 import synthetic as syn

    shapes = [(20, 31), (10, 16), (5, 8), (3, 4)]
    i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=3)
    from richsem_b200 import _capi

    args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
    win = {"_kernel": _capi.KERNEL_WINDOW}
    a = _ext().ms_deform_attn_forward(*args, 64, **win)
    ga = _ext().ms_deform_attn_backward(*args, i["grad_out"], 64, **win)
    monkeypatch.setenv("MSDA_B200_QUERY_ORDER", "natural")
    _ext()._plans.clear()  # the order is part of the cached per-signature plan
    b = _ext().ms_deform_attn_forward(*args, 64, **win)
    gb = _ext().ms_deform_attn_backward(*args, i["grad_out"], 64, **win)
    assert torch.equal(a, b)                       # forward: each (q, m) is summed in the same order
    # the order decides which levels a block serves from its shared-memory window, and the windowed
    # and direct backward paths reduce over lanes in different orders
    assert rel_err(ga[1], gb[1]) < 1e-6 and rel_err(ga[2], gb[2]) < 1e-6
    assert rel_err(ga[0], gb[0]) < 1e-5            # grad_value: atomic order differs


def test_size_independent_properties_full_size():
    """Linearity in value; constant value -> output = sum of in-range weights; zero-padding region."""
    from richsem_b200 import synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=5)
    args = (i["shapes"], i["starts"], i["loc"], i["attw"])
    f = _ext().ms_deform_attn_forward
    v1, v2 = i["value"], torch.randn_like(i["value"])
    o1, o2, o12 = f(v1, *args, 64), f(v2, *args, 64), f(v1 + 2 * v2, *args, 64)
    assert rel_err(o12, o1 + 2 * o2) < 1e-5
    ones = f(torch.ones_like(v1), *args, 64)
    assert ones.max() <= 1 + 1e-5 and ones.min() >= -1e-6
    # interior queries (all 16 samples fully inside every level) see exactly the softmax sum = 1
    assert (ones > 1 - 1e-5).float().mean() > 0.5
    # adjoint identity: <out(v), g> == <v, grad_value(g)>  (forward is linear in value)
    g = i["grad_out"]
    gv = _ext().ms_deform_attn_backward(v1, *args, g, 64)[0]
    lhs = (o1.double() * g.double()).sum()
    rhs = (v1.double() * gv.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


@pytest.mark.parametrize("kind,dtype", [("E", torch.float32), ("E", torch.bfloat16), ("U", torch.float32)])
def test_deterministic_window_backward_large_problem(kind, dtype, c_oracle):
    """Encoder-sized problems take the window kernel in deterministic mode: canonical order inside a block,
    fixed-point (order-independent) accumulation across blocks.  Bitwise reproducible, within 1e-4 of the
    oracle; "U" (no locality) sends most levels through the direct path."""
    from richsem_b200 import _capi, synthetic as syn

    shapes = [(64, 84), (32, 42), (16, 21), (8, 11)]
    i = syn.make_inputs(kind, 2, shapes, "cuda:0", seed=12, lq=None if kind == "E" else 6000, dtype=dtype)
    if kind == "U":
        i["loc"] = (i["loc"] * 1.3 - 0.15).contiguous()
    args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"], i["grad_out"], 64)
    # "U" has no query order (Lq != S): the window kernel is forced, so that its direct pass is covered too
    kw = {} if kind == "E" else {"_kernel": _capi.KERNEL_WINDOW}
    b0 = _ext().ms_deform_attn_backward
    b = lambda *a, **k: b0(*a, **k, **kw)
    d1 = b(*args, _flags=_capi.FLAG_DETERMINISTIC)
    d2 = b(*args, _flags=_capi.FLAG_DETERMINISTIC)
    at = b(*args)
    for x, y in zip(d1, d2):
        assert torch.equal(x, y)
    for x, y in zip(d1, at):
        assert rel_err(x, y) < (BWD_TOL if dtype == torch.float32 else 1e-3)
    if dtype == torch.float32:
        v, loc, w, go = (i[k].cpu() for k in ("value", "loc", "attw", "grad_out"))
        wv, wl, wa = c_oracle.backward(go, v, shapes, loc, w)
        assert rel_err(d1[0].cpu(), wv) < BWD_TOL and rel_err(d1[1].cpu(), wl) < BWD_TOL and rel_err(d1[2].cpu(), wa) < BWD_TOL
    # scale robustness: tiny and huge gradients keep their relative accuracy
    for scale in (1e-12, 1e12):
        go2 = (i["grad_out"].float() * scale).to(dtype)
        x = b(*args[:5], go2, 64, _flags=_capi.FLAG_DETERMINISTIC)[0]
        y = b(*args[:5], go2, 64)[0]
        assert rel_err(x, y) < (BWD_TOL if dtype == torch.float32 else 1e-3)


def test_deterministic_mode_is_bitwise_reproducible_and_close_to_atomic():
    from richsem_b200 import _capi, synthetic as syn

    shapes = [(40, 61), (20, 31), (10, 16), (5, 8)]
    i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=9)
    args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"], i["grad_out"], 64)
    b = _ext().ms_deform_attn_backward
    d1 = b(*args, _flags=_capi.FLAG_DETERMINISTIC)
    d2 = b(*args, _flags=_capi.FLAG_DETERMINISTIC)
    at = b(*args)
    for x, y in zip(d1, d2):
        assert torch.equal(x, y)
    for x, y in zip(d1, at):
        assert rel_err(x, y) < BWD_TOL
    # generic-kernel flavour (D != 32) and a hot spot: every query samples the same pixel
    shapes2 = [(6, 5), (3, 3)]
    j = syn.make_inputs("U", 1, shapes2, "cuda:0", seed=2, lq=3000, m=2, d=7, p=2)
    j["loc"][:, :, 0] = 0.5
    args2 = (j["value"], j["shapes"], j["starts"], j["loc"].contiguous(), j["attw"], j["grad_out"], 64)
    e1 = b(*args2, _flags=_capi.FLAG_DETERMINISTIC)
    e2 = b(*args2, _flags=_capi.FLAG_DETERMINISTIC)
    e3 = b(*args2)
    assert torch.equal(e1[0], e2[0])
    assert rel_err(e1[0], e3[0]) < BWD_TOL


def test_window_backward_matches_tiled_backward_bf16_and_five_levels():
    """Window backward vs the plain atomic one: bf16 value, and a 5-level pyramid (kL=5 decodes two
    levels in some threads)."""
    from richsem_b200 import _capi, synthetic as syn

    b = _ext().ms_deform_attn_backward
    for shapes, dtype in (([(40, 61), (20, 31), (10, 16), (5, 8)], torch.bfloat16),
                          ([(33, 47), (17, 24), (9, 12), (5, 6), (3, 3)], torch.float32)):
        i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=4, dtype=dtype)
        args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"], i["grad_out"], 64)
        y = b(*args, _kernel=_capi.KERNEL_TILED)
        x = b(*args, _kernel=_capi.KERNEL_WINDOW)
        assert rel_err(x[0], y[0]) < 1e-5
        assert rel_err(x[1], y[1]) < 1e-5 and rel_err(x[2], y[2]) < 1e-5


@pytest.mark.parametrize("path", ["auto", "window", "tiled", "generic"])
def test_bf16_value_variant(c_oracle, path):
    from richsem_b200 import _capi, synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    i = syn.make_inputs("Dn", 2, shapes, "cuda:0", seed=31, lq=1100)
    vb, gob = i["value"].bfloat16(), i["grad_out"].bfloat16()
    out = _ext().ms_deform_attn_forward(vb, i["shapes"], i["starts"], i["loc"], i["attw"], 64, **_path(path))
    assert out.dtype == torch.bfloat16
    gv, gl, ga = _ext().ms_deform_attn_backward(vb, i["shapes"], i["starts"], i["loc"], i["attw"], gob, 64,
                                                **_path(path))
    assert gv.dtype == torch.float32 and gl.dtype == torch.float32 and ga.dtype == torch.float32
    # oracle in fp32 on the UNROUNDED inputs: the 1e-2 budget covers bf16 storage of value/out/grad_out
    v, loc, w, go = (i[k].cpu() for k in ("value", "loc", "attw", "grad_out"))
    want = c_oracle.forward(v, shapes, loc, w)
    assert rel_err(out.float().cpu(), want) < BF16_TOL
    wv, wl, wa = c_oracle.backward(go, v, shapes, loc, w)
    assert rel_err(gv.cpu(), wv) < BF16_TOL
    assert rel_err(gl.cpu(), wl) < BF16_TOL
    assert rel_err(ga.cpu(), wa) < BF16_TOL


def test_empty_and_edge_inputs():
    from richsem_b200 import synthetic as syn

    shapes = [(8, 11), (4, 6), (2, 3), (1, 2)]
    shp, st, s = syn.level_tensors(shapes, "cuda:0")
    f, b = _ext().ms_deform_attn_forward, _ext().ms_deform_attn_backward
    v = torch.randn(2, s, 8, 32, device="cuda")
    # zero queries
    loc0 = torch.zeros(2, 0, 8, 4, 4, 2, device="cuda")
    w0 = torch.zeros(2, 0, 8, 4, 4, device="cuda")
    assert f(v, shp, st, loc0, w0, 64).shape == (2, 0, 256)
    gv, gl, ga = b(v, shp, st, loc0, w0, torch.zeros(2, 0, 256, device="cuda"), 64)
    assert gv.abs().max() == 0 and gl.numel() == 0
    # everything out of range -> zeros everywhere
    loc = torch.full((2, 5, 8, 4, 4, 2), 3.0, device="cuda")
    w = torch.rand(2, 5, 8, 4, 4, device="cuda")
    assert f(v, shp, st, loc, w, 64).abs().max() == 0
    gv, gl, ga = b(v, shp, st, loc, w, torch.randn(2, 5, 256, device="cuda"), 64)
    assert gv.abs().max() == 0 and gl.abs().max() == 0 and ga.abs().max() == 0
    # NaN locations are skipped (fail the range test), like the reference kernel
    loc = torch.full((2, 5, 8, 4, 4, 2), float("nan"), device="cuda")
    assert f(v, shp, st, loc, w, 64).abs().max() == 0
    # bad level table is refused, not read out of bounds
    bad = st.clone(); bad[-1] = s
    with pytest.raises(RuntimeError, match="does not fit spatial_size"):
        f(v, shp, bad, loc, w, 64)


def test_module_forward_backward_matches_oracle_path():
    """MSDeformAttn module end to end vs the same module maths with the oracle core (fp32)."""
    from oracle.msda_oracle import core_pytorch
    from richsem_b200 import synthetic as syn
    from richsem_b200.ops.modules import MSDeformAttn

    torch.manual_seed(0)
    shapes = [(16, 21), (8, 11), (4, 6), (2, 3)]
    shp, st, s = syn.level_tensors(shapes, "cuda:0")
    mod = MSDeformAttn().cuda()
    with torch.no_grad():  # move off the degenerate init so every gradient path is exercised
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    src = torch.randn(2, s, 256, device="cuda", requires_grad=True)
    ref = syn.encoder_reference_points(shapes, "cuda:0")[None, :, None, :].expand(2, s, 4, 2).contiguous()
    mask = torch.zeros(2, s, dtype=torch.bool, device="cuda"); mask[1, -5:] = True
    out = mod(src, ref, src, shp, st, mask)
    out.square().mean().backward()
    g_src = src.grad.clone(); g_off = mod.sampling_offsets.weight.grad.clone()

    # oracle path on CPU
    cpu = MSDeformAttn(); cpu.load_state_dict(mod.state_dict())
    src_c = src.detach().cpu().requires_grad_(True)
    import torch.nn.functional as F
    value = cpu.value_proj(src_c).masked_fill(mask.cpu()[..., None], 0.0).view(2, s, 8, 32)
    off = cpu.sampling_offsets(src_c).view(2, s, 8, 4, 4, 2)
    w = F.softmax(cpu.attention_weights(src_c).view(2, s, 8, 16), -1).view(2, s, 8, 4, 4)
    wh = torch.tensor([[w_, h_] for h_, w_ in shapes], dtype=torch.float32)
    loc = ref.cpu()[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    want = cpu.output_proj(core_pytorch(value, shapes, loc, w))
    want.square().mean().backward()
    assert rel_err(out.detach().cpu(), want.detach()) < 1e-4
    assert rel_err(g_src.cpu(), src_c.grad) < 1e-3
    assert rel_err(g_off.cpu(), cpu.sampling_offsets.weight.grad) < 1e-3


def test_value_without_grad_skips_grad_value_and_keeps_the_other_gradients():
    from richsem_b200 import synthetic as syn
    from richsem_b200.ops.functions import MSDeformAttnFunction

    for shapes, lq in (([(40, 61), (20, 31), (10, 16), (5, 8)], None), ([(12, 17), (6, 9)], 50)):
        i = syn.make_inputs("E" if lq is None else "U", 2, shapes, "cuda:0", seed=8, lq=lq)
        got = {}
        for need in (True, False):
            v = i["value"].clone().requires_grad_(need)
            loc = i["loc"].clone().requires_grad_(True)
            w = i["attw"].clone().requires_grad_(True)
            MSDeformAttnFunction.apply(v, i["shapes"], i["starts"], loc, w, 64).backward(i["grad_out"])
            got[need] = (v.grad, loc.grad, w.grad)
        assert got[False][0] is None and got[True][0] is not None
        assert rel_err(got[False][1], got[True][1]) < 1e-6 and rel_err(got[False][2], got[True][2]) < 1e-6


def test_forward_backward_are_cuda_graph_capturable():
    """No device synchronisation, no host read of device memory on the hot path: a forward + backward pair
    can be captured once and replayed."""
    from richsem_b200 import synthetic as syn

    shapes = [(64, 84), (32, 42), (16, 21), (8, 11)]
    i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=6)
    args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
    ext = _ext()
    want_out = ext.ms_deform_attn_forward(*args, 64)                      # also warms the host-side caches
    want = ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out = ext.ms_deform_attn_forward(*args, 64)
            grads = ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):
        out.zero_()
        for t in grads:
            t.fill_(7.0)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want_out)
        assert rel_err(grads[0], want[0]) < 1e-5          # atomic order differs between runs
        assert rel_err(grads[1], want[1]) < 1e-6 and rel_err(grads[2], want[2]) < 1e-6


@pytest.mark.parametrize("ref_dim,dtype", [(2, torch.float32), (4, torch.float32), (2, torch.bfloat16)])
def test_fused_prologue_matches_the_unfused_path(ref_dim, dtype):
    """softmax + sampling-location arithmetic inside the kernels.  The fused entry points, fed the raw
    offsets / logits / reference points, must reproduce the reference op fed the module's own PyTorch expressions
    (ms_deform_attn.py:98-111) — outputs and all gradients, through autograd for the raw tensors."""
    from richsem_b200 import synthetic as syn
    from richsem_b200.ops.functions import MSDeformAttnFunction, MSDeformAttnFusedFunction

    shapes = [(64, 84), (32, 42), (16, 21), (8, 11)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    g = torch.Generator(device=dev).manual_seed(17)
    n, m, d, L, P = 2, 8, 32, 4, 4
    value = torch.randn(n, S, m, d, generator=g, device=dev).to(dtype)
    if ref_dim == 2:
        ref = syn.encoder_reference_points(shapes, dev)[None, :, None, :].expand(n, S, L, 2).contiguous()
        offsets = (syn.head_directions(m, dev)[None, None, :, None, None, :] * torch.arange(1, P + 1, device=dev).view(1, 1, 1, 1, P, 1)
                   + 0.7 * torch.randn(n, S, m, L, P, 2, generator=g, device=dev)).contiguous()
    else:
        c = torch.rand(n, S, 1, 2, generator=g, device=dev)
        wh = torch.rand(n, S, 1, 2, generator=g, device=dev) * 0.2 + 0.02
        ref = torch.cat([c, wh], -1).expand(n, S, L, 4).contiguous()
        offsets = (2.0 * torch.randn(n, S, m, L, P, 2, generator=g, device=dev)).contiguous()
    logits = torch.randn(n, S, m, L * P, generator=g, device=dev)
    grad_out = torch.randn(n, S, m * d, generator=g, device=dev).to(dtype)

    def run(fused):
        v = value.clone().requires_grad_(True)
        off = offsets.clone().requires_grad_(True)
        lg = logits.clone().requires_grad_(True)
        if fused:
            out = MSDeformAttnFusedFunction.apply(v, shp, starts, ref, off, lg, 64)
        else:
            w = torch.softmax(lg, -1).view(n, S, m, L, P)
            if ref_dim == 2:
                norm = torch.stack([shp[..., 1], shp[..., 0]], -1)
                loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
            else:
                loc = ref[:, :, None, :, None, :2] + off / P * ref[:, :, None, :, None, 2:] * 0.5
            out = MSDeformAttnFunction.apply(v, shp, starts, loc, w, 64)
        out.backward(grad_out)
        return out.detach(), v.grad, off.grad, lg.grad

    a, b = run(True), run(False)
    tol_f, tol_b = (FWD_TOL, BWD_TOL) if dtype == torch.float32 else (1e-2, 1e-2)
    assert rel_err(a[0], b[0]) < tol_f
    for x, y in zip(a[1:], b[1:]):
        assert rel_err(x, y) < tol_b


def test_module_with_fused_prologue_matches_the_default_module():
    from richsem_b200 import synthetic as syn
    from richsem_b200.ops.modules import MSDeformAttn

    shapes = [(64, 84), (32, 42), (16, 21), (8, 11)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(5)
    plain = MSDeformAttn(256, 4, 8, 4).to(dev)
    with torch.no_grad():  # away from the all-zero initialisation, so that every parameter gets a gradient
        plain.sampling_offsets.weight.normal_(0, 0.02)
        plain.attention_weights.weight.normal_(0, 0.05)
    fused = MSDeformAttn(256, 4, 8, 4, fuse_prologue=True).to(dev)
    fused.load_state_dict(plain.state_dict())
    src = torch.randn(2, S, 256, device=dev)
    ref = syn.encoder_reference_points(shapes, dev)[None, :, None, :].expand(2, S, 4, 2).contiguous()
    mask = torch.zeros(2, S, dtype=torch.bool, device=dev)
    mask[1, -500:] = True
    outs = []
    for mod in (plain, fused):
        x = src.clone().requires_grad_(True)
        y = mod(x, ref, x, shp, starts, mask)
        y.square().mean().backward()
        outs.append((y.detach(), x.grad, [p.grad for p in mod.parameters()]))
    assert rel_err(outs[1][0], outs[0][0]) < 1e-5
    assert rel_err(outs[1][1], outs[0][1]) < 1e-4
    for gf, gp in zip(outs[1][2], outs[0][2]):
        assert rel_err(gf, gp) < 1e-4
