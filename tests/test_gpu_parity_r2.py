"""GPU parity tests added in round 2 ("parity holes"): the high-resolution shapes of BASELINE config 5
against the C oracle, full-size bitwise reproducibility of the deterministic mode, the fused prologue and the
encoder layer against the ORACLE (not against the repo's own unfused kernels), the bf16 window backward at the DINO
encoder shape, and the module under torch.inference_mode().

Tolerances as in test_gpu_parity.py: fp32 forward 1e-5, fp32 backward 1e-4, bf16 1e-2 ("relative" = max|a-b| / max|b|),
corner indices bit-exact.
"""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

FWD_TOL, BWD_TOL, BF16_TOL = 1e-5, 1e-4, 1e-2


def _ext():
    from richsem_b200 import MultiScaleDeformableAttention as ext

    return ext


@pytest.mark.parametrize("hw,S", [((1333, 1333), 37150), ((1600, 2000), 66450)])
def test_high_resolution_shapes_against_c_oracle(hw, S, c_oracle):
    """BASELINE config 5: one encoder layer at 1333x1333 / 1600x2000, N = 2, fp32, locations "E":
    forward, the three gradients and the corner indices against the C oracle."""
    from richsem_b200 import synthetic as syn

    shapes = syn.level_shapes(*hw)
    i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=51)
    assert i["S"] == S
    args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
    out = _ext().ms_deform_attn_forward(*args, 64)
    gv, gl, ga = _ext().ms_deform_attn_backward(*args, i["grad_out"], 64)
    v, loc, w, go = (i[k].cpu() for k in ("value", "loc", "attw", "grad_out"))
    assert rel_err(out.cpu(), c_oracle.forward(v, shapes, loc, w)) < FWD_TOL
    wv, wl, wa = c_oracle.backward(go, v, shapes, loc, w)
    assert rel_err(gv.cpu(), wv) < BWD_TOL
    assert rel_err(gl.cpu(), wl) < BWD_TOL
    assert rel_err(ga.cpu(), wa) < BWD_TOL
    assert torch.equal(_ext().debug_corners(i["shapes"], i["starts"], i["loc"]).cpu(), c_oracle.corners(shapes, loc))


def test_deterministic_mode_full_size_bitwise_and_close_to_atomic():
    """at S = 66,450 (N = 2) the deterministic mode is bitwise identical run to run — all three
    gradients — and within 1e-4 of the atomic mode."""
    from richsem_b200 import _capi, synthetic as syn

    shapes = syn.level_shapes(1600, 2000)
    i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=52)
    args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"], i["grad_out"], 64)
    b = _ext().ms_deform_attn_backward
    d1 = [t.clone() for t in b(*args, _flags=_capi.FLAG_DETERMINISTIC)]
    # perturb the schedule between the two runs: another kernel's blocks are resident while the second run starts
    torch.empty(1 << 28, device="cuda:0").normal_()
    d2 = b(*args, _flags=_capi.FLAG_DETERMINISTIC)
    for x, y in zip(d1, d2):
        assert torch.equal(x, y)
    at = b(*args)
    for x, y in zip(d1, at):
        assert rel_err(x, y) < BWD_TOL


def _raw_inputs(ref_dim, dtype, shapes, n, lq, seed, dev="cuda:0"):
    """Raw outputs of the module's two Linear layers (offsets, logits) + reference points, as the fused entry points take
    them.  ref_dim 2: encoder self-attention (lq must be S); 4: decoder-style boxes."""
    from richsem_b200 import synthetic as syn

    g = torch.Generator(device=dev).manual_seed(seed)
    shp, starts, S = syn.level_tensors(shapes, dev)
    m, d, L, P = 8, 32, len(shapes), 4
    value = torch.randn(n, S, m, d, generator=g, device=dev).to(dtype)
    if ref_dim == 2 and lq == S:
        ref = syn.encoder_reference_points(shapes, dev)[None, :, None, :].expand(n, S, L, 2).contiguous()
        offsets = (syn.head_directions(m, dev)[None, None, :, None, None, :] * torch.arange(1, P + 1, device=dev).view(1, 1, 1, 1, P, 1)
                   + 0.7 * torch.randn(n, lq, m, L, P, 2, generator=g, device=dev)).contiguous()
    elif ref_dim == 2:  # decoder queries with reference POINTS: anywhere in the image, offsets of a few pixels
        ref = torch.rand(n, lq, 1, 2, generator=g, device=dev).expand(n, lq, L, 2).contiguous()
        offsets = (3.0 * torch.randn(n, lq, m, L, P, 2, generator=g, device=dev)).contiguous()
    else:
        c = torch.rand(n, lq, 1, 2, generator=g, device=dev)
        wh = torch.rand(n, lq, 1, 2, generator=g, device=dev) * 0.3 + 0.02
        ref = torch.cat([c, wh], -1).expand(n, lq, L, 4).contiguous()
        offsets = (2.0 * torch.randn(n, lq, m, L, P, 2, generator=g, device=dev)).contiguous()
    logits = torch.randn(n, lq, m, L * P, generator=g, device=dev)
    grad_out = torch.randn(n, lq, m * d, generator=g, device=dev).to(dtype)
    return value, shp, starts, ref, offsets, logits, grad_out


def _oracle_prologue(c_oracle, value, shapes, ref, offsets, logits, grad_out):
    """The module's own expressions (ms_deform_attn.py:98-111) in PyTorch on the CPU around the ORACLE as the sampling
    core; autograd carries the oracle's grad_sampling_loc / grad_attn_weight back to the RAW tensors.  The core is the C
    oracle: it evaluates the pixel coordinates exactly as the CUDA contract does (cuh:285-288), so a sample that lands
    within an ulp of the pixel lattice takes the same bilinear cell in both (the grid_sample formulation rounds its
    coordinates differently there, and grad_sampling_loc is discontinuous across the lattice)."""
    n, lq, m, L, P, _ = offsets.shape
    v = value.detach().float().cpu()
    off = offsets.detach().cpu().requires_grad_(True)
    lg = logits.detach().cpu().requires_grad_(True)
    r = ref.detach().cpu()
    w = torch.softmax(lg, -1).view(n, lq, m, L, P)
    if r.shape[-1] == 2:
        norm = torch.tensor([[wd, ht] for ht, wd in shapes], dtype=torch.float32)
        loc = r[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = r[:, :, None, :, None, :2] + off / P * r[:, :, None, :, None, 2:] * 0.5
    go = grad_out.detach().float().cpu()
    out = c_oracle.forward(v, shapes, loc.detach().contiguous(), w.detach().contiguous())
    gv, gl, ga = c_oracle.backward(go, v, shapes, loc.detach().contiguous(), w.detach().contiguous())
    torch.autograd.backward([loc, w], [gl, ga])
    return out, gv, off.grad, lg.grad


@pytest.mark.parametrize("ref_dim,dtype", [(2, torch.float32), (4, torch.float32), (2, torch.bfloat16), (4, torch.bfloat16)])
def test_fused_prologue_encoder_sized_against_the_oracle(ref_dim, dtype, c_oracle):
    """SURVEY 8f-1 at encoder size (window backward, tiled forward): the fused entry points fed the raw offsets /
    logits / reference points against the reference module's PyTorch expressions + grid_sample oracle."""
    from richsem_b200.ops.functions import MSDeformAttnFusedFunction

    shapes = [(64, 84), (32, 42), (16, 21), (8, 11)]
    S = sum(h * w for h, w in shapes)
    value, shp, starts, ref, offsets, logits, grad_out = _raw_inputs(ref_dim, dtype, shapes, 2, S, seed=61)
    v = value.clone().requires_grad_(True)
    off = offsets.clone().requires_grad_(True)
    lg = logits.clone().requires_grad_(True)
    out = MSDeformAttnFusedFunction.apply(v, shp, starts, ref, off, lg, 64)
    out.backward(grad_out)
    want = _oracle_prologue(c_oracle, value, shapes, ref, offsets, logits, grad_out)
    tol_f, tol_b = (FWD_TOL, BWD_TOL) if dtype == torch.float32 else (BF16_TOL, BF16_TOL)
    assert rel_err(out.detach().float().cpu(), want[0]) < tol_f
    assert rel_err(v.grad.cpu(), want[1]) < tol_b
    assert rel_err(off.grad.cpu(), want[2]) < tol_b
    assert rel_err(lg.grad.cpu(), want[3]) < tol_b


@pytest.mark.parametrize("ref_dim,dtype", [(4, torch.float32), (2, torch.float32), (4, torch.bfloat16), (2, torch.bfloat16)])
def test_fused_prologue_decoder_sized_against_the_oracle(ref_dim, dtype, c_oracle):
    """SURVEY 8f-1 for decoder cross-attention (BASELINE config 3 shape: 900 + 200 queries, bs = 2, DINO pyramid; split
    kernels): 4-d reference boxes as at deformable_transformer.py:1017-1019 / ms_deform_attn.py:106-108, and 2-d
    reference points; against the reference module's expressions with the oracle as the sampling core."""
    from richsem_b200 import MultiScaleDeformableAttention as ext, synthetic as syn
    from richsem_b200.ops.functions import MSDeformAttnFusedFunction

    shapes = syn.level_shapes(800, 1333)
    value, shp, starts, ref, offsets, logits, grad_out = _raw_inputs(ref_dim, dtype, shapes, 2, 1100, seed=62)
    assert ext.fused_prologue_supported(value, 4, 1100, 4)
    v = value.clone().requires_grad_(True)
    off = offsets.clone().requires_grad_(True)
    lg = logits.clone().requires_grad_(True)
    out = MSDeformAttnFusedFunction.apply(v, shp, starts, ref, off, lg, 64)
    out.backward(grad_out)
    want = _oracle_prologue(c_oracle, value, shapes, ref, offsets, logits, grad_out)
    tol_f, tol_b = (FWD_TOL, BWD_TOL) if dtype == torch.float32 else (BF16_TOL, BF16_TOL)
    assert rel_err(out.detach().float().cpu(), want[0]) < tol_f
    assert rel_err(v.grad.float().cpu(), want[1]) < tol_b
    assert rel_err(off.grad.cpu(), want[2]) < tol_b
    assert rel_err(lg.grad.cpu(), want[3]) < tol_b


def test_module_with_fused_prologue_decoder_call_matches_the_default_module():
    """The module with fuse_prologue=True on a decoder-style call (4-d reference boxes, Lq = 300 != S)."""
    from richsem_b200 import synthetic as syn
    from richsem_b200.ops.modules import MSDeformAttn

    shapes = [(64, 84), (32, 42), (16, 21), (8, 11)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(6)
    plain = MSDeformAttn(256, 4, 8, 4).to(dev)
    with torch.no_grad():
        plain.sampling_offsets.weight.normal_(0, 0.05)
        plain.attention_weights.weight.normal_(0, 0.05)
    fused = MSDeformAttn(256, 4, 8, 4, fuse_prologue=True).to(dev)
    fused.load_state_dict(plain.state_dict())
    memory = torch.randn(2, S, 256, device=dev)
    query = torch.randn(2, 300, 256, device=dev)
    boxes = torch.cat([torch.rand(2, 300, 1, 2, device=dev), torch.rand(2, 300, 1, 2, device=dev) * 0.3 + 0.02], -1).expand(2, 300, 4, 4).contiguous()
    outs = []
    for mod in (plain, fused):
        q = query.clone().requires_grad_(True)
        mem = memory.clone().requires_grad_(True)
        y = mod(q, boxes, mem, shp, starts, None)
        y.square().mean().backward()
        outs.append((y.detach(), q.grad, mem.grad, [p.grad for p in mod.parameters()]))
    assert rel_err(outs[1][0], outs[0][0]) < 1e-5
    assert rel_err(outs[1][1], outs[0][1]) < 1e-4 and rel_err(outs[1][2], outs[0][2]) < 1e-4
    for gf, gp in zip(outs[1][3], outs[0][3]):
        assert rel_err(gf, gp) < 1e-4


def test_bf16_window_backward_at_the_dino_encoder_shape_against_the_oracle(c_oracle):
    """bf16 value / grad_out, N = 2, S = Lq = 22,223 (window backward): 1e-2 against the C oracle evaluated in fp32 on the
    UNROUNDED inputs (the budget covers the bf16 storage of value / out / grad_out)."""
    from richsem_b200 import synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=71)
    vb, gob = i["value"].bfloat16(), i["grad_out"].bfloat16()
    args = (vb, i["shapes"], i["starts"], i["loc"], i["attw"])
    out = _ext().ms_deform_attn_forward(*args, 64)
    gv, gl, ga = _ext().ms_deform_attn_backward(*args, gob, 64)
    v, loc, w, go = (i[k].cpu() for k in ("value", "loc", "attw", "grad_out"))
    assert rel_err(out.float().cpu(), c_oracle.forward(v, shapes, loc, w)) < BF16_TOL
    wv, wl, wa = c_oracle.backward(go, v, shapes, loc, w)
    assert rel_err(gv.cpu(), wv) < BF16_TOL
    assert rel_err(gl.cpu(), wl) < BF16_TOL
    assert rel_err(ga.cpu(), wa) < BF16_TOL
    # and tighter against the oracle fed the ROUNDED inputs: what is left is the kernel's own arithmetic
    wv2, wl2, wa2 = c_oracle.backward(gob.float().cpu(), vb.float().cpu(), shapes, loc, w)
    assert rel_err(gv.cpu(), wv2) < BWD_TOL and rel_err(gl.cpu(), wl2) < BWD_TOL and rel_err(ga.cpu(), wa2) < BWD_TOL


@pytest.mark.parametrize("padding", [False, True])
def test_encoder_layer_against_the_reference_layer_expressions(padding):
    """DeformableEncoderLayer against the reference layer's expression sequence
    (deformable_transformer.py:868-881: self_attn(with_pos_embed(src, pos), ...), residual + norm1, FFN
    linear2(relu(linear1(x))), residual + norm2; dropout 0) evaluated on the CPU with the grid_sample oracle as the
    sampling core and the module's own PyTorch expressions around it (ms_deform_attn.py:78-115)."""
    import torch.nn.functional as F

    from oracle.msda_oracle import core_pytorch
    from richsem_b200 import synthetic as syn
    from richsem_b200.encoder_layer import DeformableEncoderLayer, encoder_reference_points

    shapes = [(20, 27), (10, 14), (5, 7), (3, 4)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(13)
    layer = DeformableEncoderLayer().to(dev)
    with torch.no_grad():
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
        layer.self_attn.attention_weights.weight.normal_(0, 0.1)
    src = torch.randn(2, S, 256, device=dev)
    pos = torch.randn(2, S, 256, device=dev)
    ref = encoder_reference_points(shapes, 2, dev)
    mask = None
    if padding:
        mask = torch.zeros(2, S, dtype=torch.bool, device=dev)
        cur = 0
        for h, w in shapes:  # image 1: the right 40 % of every level is padding
            mm = torch.zeros(h, w, dtype=torch.bool)
            mm[:, int(0.6 * w):] = True
            mask[1, cur:cur + h * w] = mm.reshape(-1).to(dev)
            cur += h * w
    go = torch.randn(2, S, 256, device=dev)

    x = src.clone().requires_grad_(True)
    out = layer(x, pos, ref, shp, starts, mask)
    out.backward(go)
    got = (out.detach().cpu(), x.grad.cpu(), {k: p.grad.cpu() for k, p in layer.named_parameters()})

    import copy

    cpu = copy.deepcopy(layer).cpu()
    cpu.zero_grad()
    a = cpu.self_attn
    xs = src.detach().cpu().requires_grad_(True)
    q = xs + pos.cpu()                                                        # with_pos_embed (:863-865)
    value = a.value_proj(xs)
    if mask is not None:
        value = value.masked_fill(mask.cpu()[..., None], 0.0)                 # ms_deform_attn.py:95-96
    value = value.view(2, S, 8, 32)
    off = a.sampling_offsets(q).view(2, S, 8, 4, 4, 2)
    w = torch.softmax(a.attention_weights(q).view(2, S, 8, 16), -1).view(2, S, 8, 4, 4)
    norm = torch.tensor([[wd, ht] for ht, wd in shapes], dtype=torch.float32)
    loc = ref.cpu()[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    src2 = a.output_proj(core_pytorch(value, shapes, loc, w))
    y = cpu.norm1(xs + src2)                                                  # :871-872
    y = cpu.norm2(y + cpu.linear2(F.relu(cpu.linear1(y))))                    # :875, :857-861
    y.backward(go.cpu())
    assert rel_err(got[0], y.detach()) < 1e-5
    assert rel_err(got[1], xs.grad) < 1e-4
    for k, p in cpu.named_parameters():
        assert rel_err(got[2][k], p.grad) < 2e-4, k


def test_module_runs_under_inference_mode():
    """level tensors created under torch.inference_mode() have no version counter; the host-side level
    table cache must not read one."""
    from oracle.msda_oracle import core_pytorch
    from richsem_b200 import synthetic as syn
    from richsem_b200.ops.modules import MSDeformAttn
    from richsem_b200.ops.functions import gen_encoder_output_proposals

    shapes = [(20, 27), (10, 14), (5, 7), (3, 4)]
    dev = "cuda:0"
    torch.manual_seed(3)
    mod = MSDeformAttn().to(dev).eval()
    with torch.inference_mode():
        shp, starts, S = syn.level_tensors(shapes, dev)
        src = torch.randn(2, S, 256, device=dev)
        ref = syn.encoder_reference_points(shapes, dev)[None, :, None, :].expand(2, S, 4, 2).contiguous()
        out = mod(src, ref, src, shp, starts, None)
        om, op = gen_encoder_output_proposals(src, None, shp)
        value = mod.value_proj(src).view(2, S, 8, 32)
        off = mod.sampling_offsets(src).view(2, S, 8, 4, 4, 2)
        w = torch.softmax(mod.attention_weights(src).view(2, S, 8, 16), -1).view(2, S, 8, 4, 4)
        norm = torch.stack([shp[..., 1], shp[..., 0]], -1)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        want = mod.output_proj(core_pytorch(value.cpu(), shapes, loc.cpu(), w.cpu()).to(dev))
    assert rel_err(out, want) < 1e-5
    assert om.shape == src.shape and op.shape == (2, S, 4)


def test_graphed_step_refreshes_masks_and_rejects_other_level_tables():
    """GraphedTrainStep copies bool masks into static buffers on every call, and refuses integer tensors
    with other values (they are baked into the captured launches)."""
    from richsem_b200 import synthetic as syn
    from richsem_b200.encoder_layer import DeformableEncoder, GraphedTrainStep

    shapes = [(20, 27), (10, 14), (5, 7), (3, 4)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    torch.manual_seed(4)
    model = DeformableEncoder(2).to(dev)
    with torch.no_grad():
        for layer in model.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
    src = torch.randn(2, S, 256, device=dev)
    pos = torch.randn(2, S, 256, device=dev)
    valid = torch.ones(2, 4, 2, device=dev)
    m1 = torch.zeros(2, S, dtype=torch.bool, device=dev)
    m2 = m1.clone()
    m2[1, S // 2:] = True
    proj = torch.randn(2, S, 256, device=dev)
    loss_fn = lambda o: (o * proj).mean()  # (o.square().mean() would be ~1 whatever the input: the stack ends in a LayerNorm)
    step = GraphedTrainStep(model, loss_fn, (src, pos, shp, starts, valid, m1))
    la = step(src, pos, shp, starts, valid, m1).item()
    lb = step(src, pos, shp, starts, valid, m2).item()
    for p in model.parameters():
        p.grad = None
    want = loss_fn(model(src, pos, shp, starts, valid, m2)).item()
    assert la != lb and abs(lb - want) <= 1e-5 * abs(want)
    with pytest.raises(ValueError):
        step(src, pos, shp + 1, starts, valid, m1)


@pytest.mark.parametrize("fuse_prologue", [False, True])
def test_decoder_layer_against_the_reference_layer_expressions(fuse_prologue):
    """The caller of BASELINE config 3: DeformableDecoderLayer against the reference layer's expression sequence
    (deformable_transformer.py:965-974 self-attention, :998-1003 cross-attention with 4-d reference boxes through
    ms_deform_attn.py:98-111, :941-945 FFN; dropout 0) evaluated on the CPU with the grid_sample oracle as the sampling
    core."""
    import copy

    import torch.nn.functional as F

    from oracle.msda_oracle import core_pytorch
    from richsem_b200 import synthetic as syn
    from richsem_b200.decoder_layer import DeformableDecoderLayer

    shapes = [(20, 27), (10, 14), (5, 7), (3, 4)]
    dev = "cuda:0"
    shp, starts, S = syn.level_tensors(shapes, dev)
    nq, bs = 37, 2
    torch.manual_seed(17)
    layer = DeformableDecoderLayer(fuse_prologue=fuse_prologue).to(dev)
    with torch.no_grad():
        layer.cross_attn.sampling_offsets.weight.normal_(0, 0.02)
        layer.cross_attn.attention_weights.weight.normal_(0, 0.1)
    tgt = torch.randn(nq, bs, 256, device=dev)
    qpos = torch.randn(nq, bs, 256, device=dev)
    memory = torch.randn(S, bs, 256, device=dev)
    boxes = torch.cat([torch.rand(nq, bs, 1, 2, device=dev), torch.rand(nq, bs, 1, 2, device=dev) * 0.5 + 0.05], -1)
    boxes = boxes.expand(nq, bs, 4, 4).contiguous()
    mask = torch.zeros(bs, S, dtype=torch.bool, device=dev)
    cur = 0
    for h, w in shapes:  # image 1: the right 40 % of every level is padding
        mm = torch.zeros(h, w, dtype=torch.bool)
        mm[:, int(0.6 * w):] = True
        mask[1, cur:cur + h * w] = mm.reshape(-1).to(dev)
        cur += h * w
    go = torch.randn(nq, bs, 256, device=dev)

    x = tgt.clone().requires_grad_(True)
    mem = memory.clone().requires_grad_(True)
    out = layer(x, qpos, boxes, mem, mask, starts, shp)
    out.backward(go)
    got_params = {k: p.grad.cpu() for k, p in layer.named_parameters()}

    cpu = copy.deepcopy(layer).cpu()
    cpu.zero_grad()
    a = cpu.cross_attn
    xs = tgt.detach().cpu().requires_grad_(True)
    ms = memory.detach().cpu().requires_grad_(True)
    qp, bx = qpos.cpu(), boxes.cpu()
    q = k = xs + qp                                                            # :968
    y = cpu.norm2(xs + cpu.self_attn(q, k, xs)[0])                             # :969-971
    query = (y + qp).transpose(0, 1)                                           # :998
    ref = bx.transpose(0, 1)
    value = a.value_proj(ms.transpose(0, 1)).masked_fill(mask.cpu()[..., None], 0.0).view(bs, S, 8, 32)
    off = a.sampling_offsets(query).view(bs, nq, 8, 4, 4, 2)
    w = torch.softmax(a.attention_weights(query).view(bs, nq, 8, 16), -1).view(bs, nq, 8, 4, 4)
    loc = ref[:, :, None, :, None, :2] + off / 4 * ref[:, :, None, :, None, 2:] * 0.5   # ms_deform_attn.py:106-108
    y2 = a.output_proj(core_pytorch(value, shapes, loc, w)).transpose(0, 1)
    y = cpu.norm1(y + y2)                                                      # :1002-1003
    y = cpu.norm3(y + cpu.linear2(F.relu(cpu.linear1(y))))                     # :941-945
    y.backward(go.cpu())
    assert rel_err(out.detach().cpu(), y.detach()) < 1e-5
    assert rel_err(x.grad.cpu(), xs.grad) < 1e-4
    assert rel_err(mem.grad.cpu(), ms.grad) < 1e-4
    for kname, p in cpu.named_parameters():
        assert rel_err(got_params[kname], p.grad) < 2e-4, kname
