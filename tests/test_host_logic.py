"""CPU: host-side logic of the drop-in (validation, level-table cache, query order, shapes, module shell)."""
import math

import numpy as np
import pytest
import torch

import richsem_b200
from richsem_b200 import MultiScaleDeformableAttention as ext
from richsem_b200 import _capi, synthetic as syn
from richsem_b200.ops.functions import MSDeformAttnFunction
from richsem_b200.ops.modules import MSDeformAttn


def test_level_shapes_match_survey_table():
    assert syn.level_shapes(800, 1333) == [(100, 167), (50, 84), (25, 42), (13, 21)]
    assert syn.level_shapes(800, 1344) == [(100, 168), (50, 84), (25, 42), (13, 21)]
    assert syn.level_shapes(1333, 1333) == [(167, 167), (84, 84), (42, 42), (21, 21)]
    assert syn.level_shapes(1600, 2000) == [(200, 250), (100, 125), (50, 63), (25, 32)]
    _, starts, s = syn.level_tensors(syn.level_shapes(800, 1333), "cpu")
    assert s == 22223 and starts.tolist() == [0, 16700, 20900, 21950]


def test_algorithmic_bytes_match_survey_table():
    f, b = syn.algorithmic_bytes(2, 22223, 22223)
    assert round(f / 1e6, 2) == 159.29 and round(b / 1e6, 2) == 273.08
    assert (f + b) / (2 * 22223) == 9728
    f, b = syn.algorithmic_bytes(2, 22223, 1100, value_bytes=2, out_bytes=2)
    assert round(f / 1e6, 2) == 27.26 and round(b / 1e6, 2) == 76.15


@pytest.mark.parametrize("shapes", [[(100, 167), (50, 84), (25, 42), (13, 21)], [(8, 11), (4, 6), (2, 3), (1, 2)],
                                    [(1, 1)], [(7, 64)]])
def test_patch_order_is_a_permutation(shapes):
    _, starts, s = syn.level_tensors(shapes, "cpu")
    order = _capi.build_patch_order(tuple(shapes), tuple(starts.tolist()))
    assert order.dtype == np.int32 and order.shape == (s,)
    assert np.array_equal(np.sort(order), np.arange(s))
    # first block of a big level = one 8x8 patch
    h, w = shapes[0]
    if h >= 8 and w >= 8:
        first = order[:64]
        assert set(first // w) == set(range(8)) and set(first % w) == set(range(8))
    # padded flavour: every token exactly once, -1 elsewhere, whole 64-entry patches
    padded = _capi.build_patch_order(tuple(shapes), tuple(starts.tolist()), pad=True)
    assert padded.size % 64 == 0
    assert np.array_equal(np.sort(padded[padded >= 0]), np.arange(s))
    assert padded.min() >= -1
    for blk in padded.reshape(-1, 64):          # one block never mixes levels
        lv = {int(np.searchsorted(starts.numpy(), tkn, side="right")) for tkn in blk[blk >= 0]}
        assert len(lv) <= 1


def test_query_order_only_for_encoder_self_attention():
    shapes = [(8, 11), (4, 6)]
    shp, starts, s = syn.level_tensors(shapes, "cpu")
    meta = _capi.level_meta(shp, starts)
    assert _capi.query_order(meta, 37, "cpu") is None          # decoder: Lq != S
    assert (_capi.query_order(meta, s, "cpu") >= 0).sum() == s  # encoder


def test_level_meta_cache_tracks_identity_and_version():
    shp = torch.tensor([[4, 6], [2, 3]])
    st = torch.tensor([0, 24])
    a = _capi.level_meta(shp, st)
    assert _capi.level_meta(shp, st) is a
    shp[0, 0] = 5  # in-place edit bumps _version -> fresh read
    b = _capi.level_meta(shp, st)
    assert b is not a and b.shapes[0] == (5, 6)
    with pytest.raises(RuntimeError):
        _capi.level_meta(torch.tensor([1, 2, 3]), st)


def test_cpu_tensors_are_rejected_like_the_reference():
    value = torch.zeros(1, 30, 2, 2)
    shp, st = torch.tensor([[6, 4], [3, 2]]), torch.tensor([0, 24])
    loc, w = torch.zeros(1, 2, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ext.ms_deform_attn_forward(value, shp, st, loc, w, 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDeformAttnFunction.apply(value, shp, st, loc, w, 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        ext.ms_deform_attn_forward(value.transpose(1, 2), shp, st, loc, w, 2)


def test_dims_validation_messages():
    value = torch.zeros(3, 30, 2, 2)
    shp = torch.tensor([[6, 4], [3, 2]])
    loc, w = torch.zeros(3, 2, 2, 2, 2, 2), torch.zeros(3, 2, 2, 2, 2)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        ext._dims(value, shp, loc, w, 2)  # 3 % 2 != 0  (ms_deform_attn_cuda.cu:52)
    assert ext._dims(value, shp, loc, w, 64) == (3, 30, 2, 2, 2, 2, 2)
    with pytest.raises(RuntimeError, match="attn_weight shape"):
        ext._dims(value, shp, loc, torch.zeros(3, 2, 2, 2, 3), 64)
    with pytest.raises(RuntimeError, match="not implemented for"):
        ext._suffix(value.half(), loc, w)
    with pytest.raises(RuntimeError, match="must be torch.float32 when value is torch.bfloat16"):
        ext._suffix(value.bfloat16(), loc.double(), w.double())


def test_module_shell_matches_reference_layout_and_init():
    m = MSDeformAttn()
    assert sum(p.numel() for p in m.parameters()) == 230272
    assert list(m.state_dict()) == ["sampling_offsets.weight", "sampling_offsets.bias", "attention_weights.weight",
                                    "attention_weights.bias", "value_proj.weight", "value_proj.bias",
                                    "output_proj.weight", "output_proj.bias"]
    assert m.im2col_step == 64
    # ms_deform_attn.py:64-70: bias[m, l, p] = dir_m * (p + 1), identical for all levels
    bias = m.sampling_offsets.bias.view(8, 4, 4, 2)
    th = torch.arange(8, dtype=torch.float32) * (2.0 * math.pi / 8)
    d = torch.stack([th.cos(), th.sin()], -1)
    d = d / d.abs().max(-1, keepdim=True)[0]
    for p in range(4):
        assert torch.allclose(bias[:, :, p], (d * (p + 1))[:, None, :].expand(8, 4, 2))
    assert m.sampling_offsets.weight.abs().max() == 0 and m.attention_weights.weight.abs().max() == 0
    assert m.value_proj.bias.abs().max() == 0 and m.output_proj.bias.abs().max() == 0
    with pytest.raises(ValueError):
        MSDeformAttn(d_model=250, n_heads=8)


def test_module_forward_reaches_the_kernel_boundary_on_cpu():
    m = MSDeformAttn(d_model=16, n_levels=2, n_heads=2, n_points=2)
    shp, st = torch.tensor([[6, 4], [3, 2]]), torch.tensor([0, 24])
    q = torch.randn(1, 5, 16)
    ref = torch.rand(1, 5, 2, 2)
    src = torch.randn(1, 30, 16)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):  # no CPU fallback, by design
        m(q, ref, src, shp, st)
    with pytest.raises(ValueError, match="Last dim of reference_points"):
        m(q, torch.rand(1, 5, 2, 3), src, shp, st)
    with pytest.raises(AssertionError):
        m(q, ref, torch.randn(1, 31, 16), shp, st)


def test_install_as_reference_extension():
    import sys

    richsem_b200.install_as_reference_extension()
    import MultiScaleDeformableAttention as MSDA

    assert MSDA.ms_deform_attn_forward is ext.ms_deform_attn_forward
    assert MSDA.ms_deform_attn_backward is ext.ms_deform_attn_backward
    sys.modules.pop("MultiScaleDeformableAttention")


def test_encoder_reference_points_match_the_linspace_formulation():
    """get_reference_points restates deformable_transformer.py:512-525 (meshgrid of linspace(0.5, H-0.5, H)) with
    index arithmetic; both must give the same bits."""
    from richsem_b200.encoder_layer import get_reference_points

    shapes = [(13, 21), (7, 11), (4, 6), (2, 3)]
    vr = torch.tensor([[[1.0, 1.0]] * 4, [[0.7, 0.55], [0.72, 0.5], [0.75, 0.6], [1.0, 0.5]]])
    refs = []
    for lvl, (h, w) in enumerate(shapes):
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w), indexing="ij")
        ry = ry.reshape(-1)[None] / (vr[:, None, lvl, 1] * h)
        rx = rx.reshape(-1)[None] / (vr[:, None, lvl, 0] * w)
        refs.append(torch.stack((rx, ry), -1))
    want = torch.cat(refs, 1)[:, :, None] * vr[:, None]
    got = get_reference_points(shapes, vr)
    assert got.shape == (2, sum(h * w for h, w in shapes), 4, 2)
    assert torch.equal(got, want)


def test_add_layer_norm_host_side_selection_and_oracle_expression():
    """Residual + LayerNorm (deformable_transformer.py:871-872): what the kernels do not serve (CPU tensors, widths that
    are not a multiple of 128) takes the PyTorch expression; the kernel entry point itself refuses CPU tensors; the
    oracle's fp64 expression agrees with nn.LayerNorm."""
    from oracle import aux_oracle as ao
    from richsem_b200.ops.functions import add_layer_norm
    from richsem_b200.ops.functions.aux_functions import AddLayerNormFunction, add_layer_norm_supported

    g = torch.Generator().manual_seed(9)
    x, r = torch.randn(2, 5, 256, generator=g), torch.randn(2, 5, 256, generator=g)
    norm = torch.nn.LayerNorm(256)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.1, 0.1)
    assert not add_layer_norm_supported(x, r)  # CPU
    assert torch.equal(add_layer_norm(x, r, norm), norm(x + r))
    assert torch.equal(add_layer_norm(x, None, weight=norm.weight, bias=norm.bias, eps=norm.eps), norm(x))
    with pytest.raises(RuntimeError, match="CPU"):
        AddLayerNormFunction.apply(x, r, norm.weight, norm.bias, 1e-5)
    want = ao.add_layer_norm(x, r, norm.weight, norm.bias, norm.eps)
    assert (want - norm(x + r).double()).abs().max() < 1e-5
    # the C ABI's argument checks need no GPU
    lib = richsem_b200._capi.lib
    assert lib.msda_add_layernorm_workspace_bytes(44446, 256) >= 2 * 256 * 4
    assert lib.msda_add_layernorm_f32(None, None, None, None, None, 10, 96, 1e-5, None, None, None) == 2  # UNSUPPORTED
    assert b"multiple of 128" in lib.msda_last_error()
    assert lib.msda_add_layernorm_f32(None, None, None, None, None, 0, 256, 1e-5, None, None, None) == 0  # empty: nothing to do


def test_decoder_layer_parameter_names_follow_the_reference_layer():
    """deformable_transformer.py:896-917: cross_attn / norm1 / self_attn / norm2 / linear1 / linear2 / norm3 (dropouts
    have no parameters), so a reference decoder-layer state dict loads."""
    from richsem_b200.decoder_layer import DeformableDecoderLayer

    names = {k for k, _ in DeformableDecoderLayer().named_parameters()}
    want = {f"cross_attn.{m}.{p}" for m in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
            for p in ("weight", "bias")}
    want |= {"self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias"}
    want |= {f"{m}.{p}" for m in ("norm1", "norm2", "norm3", "linear1", "linear2") for p in ("weight", "bias")}
    assert names == want
