"""CPU: pins the oracle (both restatements) to the reference's own function and golden vectors."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden, rel_err
from oracle import msda_oracle as om


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_torch_restatement_reproduces_golden_bitwise(case):
    g = load_golden(case)
    out, gv, gl, ga = om.core_pytorch_fwd_bwd(g["value"], g["shape_list"], g["loc"], g["attw"], g["grad_out"])
    # same op sequence on the same torch build -> identical bits
    assert torch.equal(out, g["out"])
    assert torch.equal(gv, g["grad_value"])
    assert torch.equal(gl, g["grad_loc"])
    assert torch.equal(ga, g["grad_attw"])


def test_torch_restatement_equals_reference_function_when_present():
    ref = om.load_reference_core()
    if ref is None:
        pytest.skip("/root/reference not present (GPU box)")
    gen = torch.Generator().manual_seed(5)
    shapes = [(9, 13), (5, 7), (3, 4)]
    s = sum(h * w for h, w in shapes)
    for dtype in (torch.float32, torch.float64):
        value = torch.randn(2, s, 4, 8, generator=gen, dtype=dtype)
        loc = torch.rand(2, 19, 4, 3, 4, 2, generator=gen, dtype=dtype) * 1.4 - 0.2
        attw = torch.rand(2, 19, 4, 3, 4, generator=gen, dtype=dtype)
        a = om.core_pytorch(value, shapes, loc, attw)
        b = ref(value, torch.as_tensor(shapes), loc, attw)
        assert torch.equal(a, b)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_c_restatement_matches_golden(case, c_oracle):
    g = load_golden(case)
    f64 = g["value"].dtype == torch.float64
    tol_f, tol_b = (1e-12, 1e-11) if f64 else (1e-5, 1e-4)
    out = c_oracle.forward(g["value"], g["shape_list"], g["loc"], g["attw"])
    assert rel_err(out, g["out"]) < tol_f
    gv, gl, ga = c_oracle.backward(g["grad_out"], g["value"], g["shape_list"], g["loc"], g["attw"])
    assert rel_err(gv, g["grad_value"]) < tol_b
    assert rel_err(gl, g["grad_loc"]) < tol_b
    assert rel_err(ga, g["grad_attw"]) < tol_b


def test_c_oracle_is_deterministic(c_oracle):
    g = load_golden("enc_small_f32")
    a = c_oracle.backward(g["grad_out"], g["value"], g["shape_list"], g["loc"], g["attw"])
    b = c_oracle.backward(g["grad_out"], g["value"], g["shape_list"], g["loc"], g["attw"])
    assert all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("case", ["enc_small_f32", "pad_small_f32", "dec_small_f32", "tiny_f32"])
def test_corner_indices_c_vs_numpy(case, c_oracle):
    g = load_golden(case)
    c = c_oracle.corners(g["shape_list"], g["loc"]).numpy()
    n = om.corners_numpy(g["shape_list"], g["loc"].numpy())
    assert np.array_equal(c, n)


def test_fma_coordinate_variant_differs_only_on_the_lattice(c_oracle):
    """The compiled reference kernel contracts loc*size-0.5 into one FMA; the two
    variants may only disagree where fl(loc*size) lands on k+0.5."""
    from richsem_b200 import synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    gen = torch.Generator().manual_seed(7)
    loc = syn.locations_encoder(1, shapes, gen, "cpu", jitter_px=0.0)[:, ::5].contiguous()  # on the lattice
    a = c_oracle.corners(shapes, loc)
    b = c_oracle.corners(shapes, loc, fma=True)
    assert np.array_equal(b.numpy(), om.corners_numpy(shapes, loc.numpy(), fma=True))
    differ = (a != b).any(-1)
    assert 0 < differ.float().mean() < 0.5          # a sizeable share of lattice points flips ...
    jit = syn.locations_encoder(1, shapes, gen, "cpu", jitter_px=0.5)[:, ::5].contiguous()
    d2 = (c_oracle.corners(shapes, jit) != c_oracle.corners(shapes, jit, fma=True)).any(-1)
    assert d2.float().mean() < 1e-5                 # ... and essentially none off the lattice


def test_corner_indices_known_answers(c_oracle):
    # one level 4x6 (H=4, W=6), start 10; hand-computed from cuh:285-288 / :38-78
    shapes, start = [(4, 6)], [10]
    pts = torch.tensor([
        [0.5, 0.5],      # w_im=2.5 h_im=1.5 -> (1,2) (1,3) (2,2) (2,3)
        [0.0, 0.0],      # -0.5,-0.5 -> only (0,0)
        [1.0, 1.0],      # 5.5, 3.5 -> only (3,5)
        [-0.1, 0.5],     # w_im=-1.1 -> skipped
        [1.2, 0.5],      # w_im=6.7 >= W -> skipped
        [0.25, 0.375],   # w_im=1.0 h_im=1.0 exactly on the lattice -> floor = itself
    ], dtype=torch.float32).view(1, 1, 1, 1, 6, 2)
    c = c_oracle.corners(shapes, pts, level_start_index=start).view(6, 4).tolist()
    t = lambda h, w: 10 + h * 6 + w
    assert c[0] == [t(1, 2), t(1, 3), t(2, 2), t(2, 3)]
    assert c[1] == [-1, -1, -1, t(0, 0)]
    assert c[2] == [t(3, 5), -1, -1, -1]
    assert c[3] == [-1, -1, -1, -1]
    assert c[4] == [-1, -1, -1, -1]
    assert c[5] == [t(1, 1), t(1, 2), t(2, 1), t(2, 2)]


def test_lattice_kink_is_resolved_by_the_cuda_formula_not_grid_sample(c_oracle):
    """on the pixel lattice the contract is floor(fl(fl(x*W) - 0.5)), evaluated in fp32."""
    w = 167
    xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w + 1.0 / w  # centre + one pixel: lands on k + 1.0
    loc = torch.stack([xs, torch.full_like(xs, 0.5)], -1).view(1, w, 1, 1, 1, 2)
    c = c_oracle.corners([(100, w)], loc).view(w, 4)
    wim = (xs * np.float32(w)).float() - 0.5
    h0 = 49  # 0.5*100-0.5 = 49.5
    expect_w0 = torch.floor(wim).long()
    ok = expect_w0 <= w - 1
    got_w0 = (c[:, 0].long() - h0 * w)
    assert torch.equal(got_w0[ok & (c[:, 0] >= 0)], expect_w0[ok & (c[:, 0] >= 0)])
