"""GPU parity of the elementwise passes either side of the sampling core (kernels in
richsem_b200/csrc/msda_aux.cu) against oracle/aux_oracle.py and the golden vectors generated from the reference."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import aux_oracle as ao

pytestmark = pytest.mark.gpu

DINO = [(100, 167), (50, 84), (25, 42), (13, 21)]


def _rect(shapes, frac):
    parts = []
    for h, w in shapes:
        m = torch.ones(h, w, dtype=torch.bool)
        m[: max(1, round(h * frac[0])), : max(1, round(w * frac[1]))] = False
        parts.append(m.reshape(-1))
    return torch.cat(parts)


# ---- value preparation (8f-2) -----------------------------------------------------------------------
@pytest.mark.parametrize("rows_shape,c", [((2, 22223), 256), ((3, 41), 12), ((1, 1), 4)])
@pytest.mark.parametrize("with_mask", [True, False])
def test_value_prepare_bf16_is_bit_exact(rows_shape, c, with_mask):
    from richsem_b200.ops.functions.aux_functions import cast_value_bf16

    g = torch.Generator().manual_seed(3)
    x = torch.randn(*rows_shape, c, generator=g) * 3
    x.view(-1)[::7] *= 1e-3
    mask = (torch.rand(*rows_shape, generator=g) < 0.3) if with_mask else None
    want = ao.value_prepare(x, mask, torch.bfloat16)
    got = cast_value_bf16(x.cuda(), None if mask is None else mask.cuda()).cpu()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))


def test_zero_masked_rows_in_place_touches_only_masked_rows():
    from richsem_b200.ops.functions.aux_functions import zero_masked_rows_

    g = torch.Generator().manual_seed(4)
    for shape in ((2, 22223, 256), (5, 33, 8)):
        x = torch.randn(*shape, generator=g)
        mask = torch.rand(*shape[:-1], generator=g) < 0.4
        d = x.cuda()
        out = zero_masked_rows_(d, mask.cuda())
        assert out.data_ptr() == d.data_ptr()
        assert torch.equal(d.cpu(), ao.value_prepare(x, mask))
    # NaN / inf in a masked row are overwritten, not multiplied
    x = torch.full((1, 3, 4), float("nan"))
    d = zero_masked_rows_(x.cuda(), torch.tensor([[True, False, True]]).cuda()).cpu()
    assert torch.equal(d[0, 0], torch.zeros(4)) and torch.isnan(d[0, 1]).all()


@pytest.mark.parametrize("dtype", [None, torch.bfloat16])
def test_prepare_value_autograd_matches_masked_fill(dtype):
    from richsem_b200.ops.functions import prepare_value

    g = torch.Generator().manual_seed(5)
    lin = torch.nn.Linear(16, 32).cuda()
    x = torch.randn(2, 50, 16, generator=g).cuda().requires_grad_(True)
    mask = (torch.rand(2, 50, generator=g) < 0.3).cuda()
    go = torch.randn(2, 50, 32, generator=g).cuda()

    ref = lin(x).masked_fill(mask[..., None], 0.0)
    ref = ref if dtype is None else ref.to(dtype)
    ref.backward(go.to(ref.dtype))
    want = (ref.detach().clone(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None
    lin.zero_grad()

    out = prepare_value(lin(x), mask, dtype)
    out.backward(go.to(out.dtype))
    assert out.dtype == ref.dtype and torch.equal(out.detach(), want[0])
    assert torch.equal(x.grad, want[1]) and torch.equal(lin.weight.grad, want[2]) and torch.equal(lin.bias.grad, want[3])


@pytest.mark.parametrize("value_dtype", [None, torch.bfloat16])
def test_module_with_padding_mask_matches_reference_formulation(value_dtype):
    """MSDeformAttn with a padding mask (ms_deform_attn.py:94-97 through the new pass) against the same module
    evaluated with PyTorch's masked_fill and the grid_sample oracle."""
    from oracle.msda_oracle import core_pytorch
    from richsem_b200 import synthetic as syn
    from richsem_b200.ops.modules import MSDeformAttn

    shapes = [(20, 27), (10, 14), (5, 7), (3, 4)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(9)
    mod = MSDeformAttn(value_dtype=value_dtype).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    src = torch.randn(2, S, 256, device=dev, requires_grad=True)
    ref_pts = syn.encoder_reference_points(shapes, dev)[None, :, None, :].expand(2, S, 4, 2).contiguous()
    mask = torch.stack([_rect(shapes, (1.0, 1.0)), _rect(shapes, (0.7, 0.55))]).to(dev)
    go = torch.randn(2, S, 256, device=dev)

    out = mod(src, ref_pts, src, shp, starts, mask)
    out.backward(go)
    got = (out.detach().clone(), src.grad.clone(), mod.value_proj.weight.grad.clone())
    src.grad = None
    mod.zero_grad()

    # the reference module's expressions (ms_deform_attn.py:94-114) with the oracle as the sampling core
    value = mod.value_proj(src).masked_fill(mask[..., None], 0.0)
    if value_dtype is not None:
        value = value.to(value_dtype).float()   # same storage rounding, fp32 arithmetic
    value = value.view(2, S, 8, 32)
    off = mod.sampling_offsets(src).view(2, S, 8, 4, 4, 2)
    w = torch.softmax(mod.attention_weights(src).view(2, S, 8, 16), -1).view(2, S, 8, 4, 4)
    wh = torch.stack([shp[..., 1], shp[..., 0]], -1)
    loc = ref_pts[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    sampled = core_pytorch(value, shapes, loc, w)
    if value_dtype is not None:
        sampled = sampled.to(value_dtype).float()
    want = mod.output_proj(sampled)
    want.backward(go)
    tol_f, tol_b = (1e-5, 1e-4) if value_dtype is None else (1e-2, 1e-2)
    assert rel_err(got[0], want) < tol_f
    assert rel_err(got[1], src.grad) < tol_b
    assert rel_err(got[2], mod.value_proj.weight.grad) < tol_b
    # padded tokens get no gradient through the value path: check directly on value_proj's bias gradient
    assert torch.isfinite(got[1]).all()


# ---- two-stage proposals (8f-4) ---------------------------------------------------------------------
def _check_proposals(memory, mask, shapes, wh):
    from richsem_b200.ops.functions import gen_encoder_output_proposals

    want_m, want_p = ao.encoder_proposals(memory, mask, shapes, wh)
    dev = "cuda:0"
    shp = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    mem = memory.to(dev).requires_grad_(True)
    om, op = gen_encoder_output_proposals(mem, None if mask is None else mask.to(dev), shp,
                                          None if wh is None else wh.to(dev))
    # which tokens survive is index work: bit-exact; so is the masked copy
    assert torch.equal(torch.isinf(op).cpu(), torch.isinf(want_p))
    assert torch.equal(om.detach().cpu(), want_m)
    fin = torch.isfinite(want_p)
    assert (op.detach().cpu()[~fin] == float("inf")).all()
    # logits: same fp32 operations; logf may differ from the CPU's log in the last ulp
    err = ((op.detach().cpu()[fin].double() - want_p[fin].double()).abs() / want_p[fin].double().abs().clamp_min(1e-3)).max()
    assert err < 1e-6, err
    go = torch.randn_like(om)
    om.backward(go)
    keep = fin[..., :1].to(dev)
    assert torch.equal(mem.grad, torch.where(keep, go, torch.zeros_like(go)))
    return op


@pytest.mark.parametrize("case", ["proposals_pad", "proposals_learned"])
def test_proposals_golden(case):
    z = np.load(GOLDEN / f"{case}.npz")
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    wh = g["learnedwh"] if g["learnedwh"].numel() else None
    op = _check_proposals(g["memory"], g["mask"], [tuple(int(x) for x in r) for r in g["shapes"]], wh)
    fin = torch.isfinite(g["output_proposals"])
    assert rel_err(op.detach().cpu()[fin], g["output_proposals"][fin]) < 1e-6


@pytest.mark.parametrize("frac", [(1.0, 1.0), (0.83, 0.61), (0.08, 0.5), None])
def test_proposals_dino_shape(frac):
    g = torch.Generator().manual_seed(21)
    s = sum(h * w for h, w in DINO)
    memory = torch.randn(2, s, 256, generator=g)
    mask = None if frac is None else torch.stack([_rect(DINO, (1.0, 1.0)), _rect(DINO, frac)])
    _check_proposals(memory, mask, DINO, None)


def test_proposals_large_batch_uses_the_workspace_table():
    shapes = [(6, 9), (3, 5), (2, 2)]
    g = torch.Generator().manual_seed(8)
    n = 50   # 150 (image, level) pairs > 128: valid H / W come from the stand-alone pass
    memory = torch.randn(n, 73, 8, generator=g)
    mask = torch.stack([_rect(shapes, (1.0, 1.0) if i % 3 == 0 else (0.3 + 0.01 * i, 0.9 - 0.01 * i)) for i in range(n)])
    _check_proposals(memory, mask, shapes, None)


def test_proposals_fully_padded_image_and_errors():
    from richsem_b200.ops.functions import gen_encoder_output_proposals

    shapes = [(6, 9), (3, 5)]
    s = 69
    memory = torch.randn(2, s, 8)
    mask = torch.zeros(2, s, dtype=torch.bool)
    mask[1] = True  # valid_W = valid_H = 0: division by zero -> inf -> invalid everywhere
    _check_proposals(memory, mask, shapes, None)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        gen_encoder_output_proposals(memory, mask, torch.as_tensor(shapes), None)
    with pytest.raises(RuntimeError, match="spatial_size"):
        gen_encoder_output_proposals(memory.cuda(), mask.cuda(), torch.as_tensor([(6, 9), (3, 4)]).cuda(), None)
    wh = torch.zeros(2, device="cuda", requires_grad=True)
    with pytest.raises(NotImplementedError):
        gen_encoder_output_proposals(memory.cuda(), mask.cuda(), torch.as_tensor(shapes).cuda(), wh)


# ---- encoder stack + CUDA graph (8f-3) --------------------------------------------------------------
@pytest.mark.parametrize("padding", [False, True])
def test_graphed_encoder_step_matches_eager(padding):
    """The 6-layer encoder's forward + backward captured in one CUDA graph reproduces the eager step (same kernels;
    grad_value atomics reorder between runs, hence a tolerance), and new inputs flow through the static buffers."""
    from richsem_b200 import synthetic as syn
    from richsem_b200.encoder_layer import DeformableEncoder, GraphedTrainStep

    shapes = [(40, 54), (20, 27), (10, 14), (5, 7)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(3)
    model = DeformableEncoder(3).to(dev)
    with torch.no_grad():
        for layer in model.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    valid = torch.ones(2, 4, 2, device=dev)
    mask = None
    if padding:
        mask = torch.stack([_rect(shapes, (1.0, 1.0)), _rect(shapes, (0.7, 0.55))]).to(dev)
        valid[1, :, 0], valid[1, :, 1] = 0.55, 0.7
    loss_fn = lambda out: out.square().mean()
    src0, pos = torch.randn(2, S, 256, device=dev), torch.randn(2, S, 256, device=dev)
    step = GraphedTrainStep(model, loss_fn, (src0, pos, shp, starts, valid, mask))
    for seed in (1, 2):
        src = torch.randn(2, S, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
        loss = step(src, pos, shp, starts, valid, mask).clone()
        got = [p.grad.clone() for p in model.parameters()]
        for p in model.parameters():
            p.grad = None
        want_loss = loss_fn(model(src, pos, shp, starts, valid, mask))
        want_loss.backward()
        assert abs(loss.item() - want_loss.item()) <= 1e-5 * abs(want_loss.item())
        for g, p in zip(got, model.parameters()):
            assert rel_err(g, p.grad) < 1e-4
        for p in model.parameters():
            p.grad = None


# ---- two-stage query selection (8f-4, second half) ------------------------------------------------------
@pytest.mark.parametrize("rows_shape,k_classes", [((2, 22223), 91), ((2, 22223), 1203), ((3, 50), 1), ((1, 7), 256), ((2, 300), 31)])
def test_class_scores_match_torch_max(rows_shape, k_classes):
    from richsem_b200.ops.functions.aux_functions import class_scores

    g = torch.Generator().manual_seed(31)
    x = torch.randn(*rows_shape, k_classes, generator=g)
    got = class_scores(x.cuda()).cpu()
    assert torch.equal(got, x.max(-1)[0])          # a maximum is exact


def test_class_scores_propagate_nan_like_torch():
    from richsem_b200.ops.functions.aux_functions import class_scores

    x = torch.randn(4, 300)
    x[1, 17] = float("nan")
    x[2, 299] = float("nan")
    x[3, 5] = float("inf")
    got = class_scores(x.cuda()).cpu()
    want = x.max(-1)[0]
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(got[~torch.isnan(want)], want[~torch.isnan(want)])


@pytest.mark.parametrize("n,s,k", [(2, 22223, 900), (2, 22323, 900), (3, 1000, 1000), (1, 5000, 1), (4, 1024, 37), (2, 66450, 1024)])
def test_topk_rows_matches_torch_topk(n, s, k):
    from richsem_b200.ops.functions.aux_functions import topk_rows

    g = torch.Generator().manual_seed(41)
    scores = torch.randn(n, s, generator=g)
    val, idx = topk_rows(scores.cuda(), k, return_values=True)
    want_val, want_idx = torch.topk(scores, k, dim=1)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (n, k)
    assert torch.equal(val.cpu(), want_val)                       # the selected scores, in order: bit-exact
    assert torch.equal(scores.gather(1, idx.cpu()), want_val)     # indices point at them
    # indices are bit-exact wherever the score is unique in its row (torch leaves the order of ties unspecified)
    for b in range(n):
        v = want_val[b]
        unique = torch.ones(k, dtype=torch.bool)
        unique[1:] &= v[1:] != v[:-1]
        unique[:-1] &= v[:-1] != v[1:]
        kth = v[-1]
        unique &= ~((v == kth) & ((scores[b] == kth).sum() > (v == kth).sum()))
        assert torch.equal(idx.cpu()[b][unique], want_idx[b][unique])


def test_topk_rows_ties_nan_and_masked_tokens():
    """Ties are broken towards the lower index; NaN ranks first (torch.topk); a row of mostly equal scores — what the
    zeroed memory rows of padded tokens produce (utils.py:54-56) — still yields k distinct indices."""
    from richsem_b200.ops.functions.aux_functions import topk_rows

    s = torch.zeros(2, 5000)
    s[0, 100:110] = 1.0
    s[0, 4000] = 2.0
    s[0, 77] = float("nan")
    s[1] = -3.0
    s[1, 4999] = float("inf")
    s[1, 0] = float("-inf")
    idx = topk_rows(s.cuda(), 20).cpu()
    assert idx[0].tolist() == [77, 4000] + list(range(100, 110)) + list(range(0, 8))
    assert idx[1].tolist() == [4999] + list(range(1, 20))
    for b in range(2):
        assert len(set(idx[b].tolist())) == 20


def test_topk_proposals_is_the_reference_expression():
    from richsem_b200.ops.functions import topk_proposals

    g = torch.Generator().manual_seed(43)
    logits = torch.randn(2, 22223, 91, generator=g)
    got = topk_proposals(logits.cuda(), 900).cpu()
    want = torch.topk(logits.max(-1)[0], 900, dim=1)[1]           # deformable_transformer.py:367-369
    assert torch.equal(logits.max(-1)[0].gather(1, got), logits.max(-1)[0].gather(1, want))
    assert (got == want).float().mean() > 0.999
    with pytest.raises(RuntimeError, match="out of range"):
        topk_proposals(logits.cuda(), 30000)
    with pytest.raises(RuntimeError, match="at most 1024"):
        topk_proposals(logits.cuda(), 2000)


def test_topk_single_block_kernel_matches_the_cluster_kernel():
    """Rows too long for the cluster's shared-memory slices (> 98,304 scores) take the one-block-per-image kernel;
    both must agree with torch.topk."""
    from richsem_b200.ops.functions.aux_functions import topk_rows

    g = torch.Generator().manual_seed(47)
    for s_len in (98304, 98305, 120000):
        scores = torch.randn(2, s_len, generator=g)
        val, idx = topk_rows(scores.cuda(), 300, return_values=True)
        want_val, _ = torch.topk(scores, 300, dim=1)
        assert torch.equal(val.cpu(), want_val)
        assert torch.equal(scores.gather(1, idx.cpu()), want_val)


# ---- randomized sweeps (seeded) over shapes the fixed cases do not hit ---------------------------------------
def test_randomized_proposals_and_selection_sweep():
    """30 random pyramids (1-5 levels, odd sizes, channel counts that are any multiple of 4), random rectangular or
    scattered padding, random k: proposals against the oracle, selection against torch."""
    import random

    from richsem_b200.ops.functions import topk_proposals

    rnd = random.Random(2024)
    for it in range(30):
        levels = rnd.randint(1, 5)
        shapes = [(rnd.randint(1, 40), rnd.randint(1, 40)) for _ in range(levels)]
        s = sum(h * w for h, w in shapes)
        n = rnd.randint(1, 4)
        c = 4 * rnd.randint(1, 70)
        g = torch.Generator().manual_seed(1000 + it)
        memory = torch.randn(n, s, c, generator=g)
        kind = rnd.choice(["none", "rect", "scattered"])
        if kind == "none":
            mask = None
        elif kind == "rect":
            mask = torch.stack([_rect(shapes, (rnd.uniform(0.05, 1.0), rnd.uniform(0.05, 1.0))) for _ in range(n)])
        else:
            mask = torch.rand(n, s, generator=g) < rnd.uniform(0.0, 0.9)
        wh = None if rnd.random() < 0.7 else torch.tensor([rnd.uniform(-3, 1), rnd.uniform(-3, 1)])
        _check_proposals(memory, mask, shapes, wh)
        kc = rnd.choice([1, 3, 17, 91, 130, 515, 1203])
        logits = torch.randn(n, s, kc, generator=g)
        if rnd.random() < 0.3:  # a few masked-looking rows: identical logits
            logits[:, : s // 3] = -1.5
        k = rnd.randint(1, min(s, 1024))
        got = topk_proposals(logits.cuda(), k).cpu()
        sc = logits.max(-1)[0]
        want_val = torch.topk(sc, k, dim=1)[0]
        assert torch.equal(sc.gather(1, got), want_val), (it, shapes, kc, k)
        for b in range(n):
            assert len(set(got[b].tolist())) == k


# ---- encoder-layer epilogue: residual add + LayerNorm (8f-3) ------------------------------------------------------
@pytest.mark.parametrize("rows_shape,c", [((2, 22223), 256), ((3, 41), 128), ((1, 1), 512), ((5, 7), 384)])
@pytest.mark.parametrize("with_residual", [True, False])
def test_add_layer_norm_forward_and_backward_against_the_oracle(rows_shape, c, with_residual):
    """out and all gradients against the fp64 expression of oracle/aux_oracle.py (autograd on the CPU); torch's own fp32
    LayerNorm kernel on the same inputs sets the scale of an acceptable fp32 error."""
    from richsem_b200.ops.functions.aux_functions import AddLayerNormFunction

    g = torch.Generator().manual_seed(5)
    x = torch.randn(*rows_shape, c, generator=g) * 2 + 0.3
    r = torch.randn(*rows_shape, c, generator=g) if with_residual else None
    w = torch.randn(c, generator=g) * 0.5 + 1
    b = torch.randn(c, generator=g) * 0.1
    go = torch.randn(*rows_shape, c, generator=g)

    leaves = [t.double().requires_grad_(True) if t is not None else None for t in (x, r, w, b)]
    want = ao.add_layer_norm(*leaves, eps=1e-5)
    want.backward(go.double())

    dl = [t.cuda().requires_grad_(True) if t is not None else None for t in (x, r, w, b)]
    got = AddLayerNormFunction.apply(*dl, 1e-5)
    got.backward(go.cuda())
    assert rel_err(got.detach().cpu(), want.detach()) < 2e-6
    for name, a, e in zip(("x", "residual", "weight", "bias"), dl, leaves):
        if a is None:
            continue
        # parameter gradients are sums over up to 44,446 rows: fp32 accumulation error grows with sqrt(rows)
        tol = 2e-5 if name in ("weight", "bias") else 5e-6
        assert rel_err(a.grad.cpu(), e.grad) < tol, name
    # second witness: torch's fp32 kernel
    y = x.cuda() if r is None else x.cuda() + r.cuda()
    torch_out = torch.nn.functional.layer_norm(y, (c,), w.cuda(), b.cuda(), 1e-5)
    assert rel_err(got.detach(), torch_out) < 2e-6


def test_add_layer_norm_parameter_gradients_are_bitwise_reproducible_and_edges():
    from richsem_b200.ops.functions import add_layer_norm
    from richsem_b200.ops.functions.aux_functions import AddLayerNormFunction, add_layer_norm_supported

    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 22223, 256, generator=g).cuda()
    r = torch.randn(2, 22223, 256, generator=g).cuda()
    go = torch.randn(2, 22223, 256, generator=g).cuda()
    norm = torch.nn.LayerNorm(256).cuda()
    runs = []
    for _ in range(3):
        norm.zero_grad()
        xx = x.clone().requires_grad_(True)
        add_layer_norm(xx, r, norm).backward(go)
        runs.append((norm.weight.grad.clone(), norm.bias.grad.clone(), xx.grad.clone()))
    for k in range(3):
        assert torch.equal(runs[0][k], runs[1][k]) and torch.equal(runs[0][k], runs[2][k])
    # no parameters / no gradient wanted for them; empty input; unsupported widths take the PyTorch expression
    out = AddLayerNormFunction.apply(x[:, :9], r[:, :9], None, None, 1e-5)
    assert rel_err(out, torch.nn.functional.layer_norm(x[:, :9] + r[:, :9], (256,))) < 2e-6
    assert AddLayerNormFunction.apply(x[:, :0], r[:, :0], norm.weight, norm.bias, 1e-5).shape == (2, 0, 256)
    z = torch.randn(4, 6, 96, device="cuda")
    assert not add_layer_norm_supported(z)
    n96 = torch.nn.LayerNorm(96).cuda()
    assert torch.equal(add_layer_norm(z, z, n96), n96(z + z))
    with pytest.raises(RuntimeError):
        AddLayerNormFunction.apply(z, z, None, None, 1e-5)
    with pytest.raises(RuntimeError):
        AddLayerNormFunction.apply(x.cpu(), None, None, None, 1e-5)


def test_encoder_layer_with_fused_epilogue_under_inference_mode_and_switch():
    """The layer's one-pass residual + LayerNorm runs under torch.inference_mode() (eval setup) and gives what the
    two-kernel PyTorch expression gives (fuse_epilogue=False) on the same parameters."""
    from richsem_b200 import synthetic as syn
    from richsem_b200.encoder_layer import DeformableEncoderLayer, encoder_reference_points

    shapes = [(20, 27), (10, 14), (5, 7), (3, 4)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(21)
    layer = DeformableEncoderLayer().to(dev).eval()
    with torch.no_grad():
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
        layer.norm1.weight.uniform_(0.5, 1.5)
        layer.norm2.bias.uniform_(-0.2, 0.2)
    src, pos = torch.randn(2, S, 256, device=dev), torch.randn(2, S, 256, device=dev)
    ref = encoder_reference_points(shapes, 2, dev)
    with torch.inference_mode():
        fused = layer(src, pos, ref, shp, starts, None)
    layer.fuse_epilogue = False
    with torch.no_grad():
        plain = layer(src, pos, ref, shp, starts, None)
    assert rel_err(fused, plain) < 5e-6
