"""GPU: our kernels against the REFERENCE's own CUDA kernels (oracle/_ref, compiled unmodified from
/root/reference for sm_100a by oracle/build_ref.py) on identical inputs."""
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def legacy():
    from oracle.msda_oracle import LegacyCuda

    if not LegacyCuda.available():
        pytest.skip("oracle/_ref/libmsda_legacy.so not built (needs /root/reference at build time)")
    return LegacyCuda()


@pytest.mark.parametrize("kind,n,lq", [("E", 2, None), ("U", 1, 3000), ("Dn", 2, 1100)])
def test_dino_shape_matches_reference_cuda_kernels(legacy, kind, n, lq):
    from richsem_b200 import MultiScaleDeformableAttention as ext, synthetic as syn

    shapes = syn.level_shapes(800, 1333)
    i = syn.make_inputs(kind, n, shapes, "cuda:0", seed=77, lq=lq)
    from richsem_b200 import _capi

    args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
    ref_out = legacy.forward(*args)
    ref_gv, ref_gl, ref_ga = legacy.backward(*args, i["grad_out"])
    # (a) as-compiled arithmetic: nvcc contracts the reference's loc*size-0.5 into one FMA, and with
    #     MSDA_FLAG_COORDS_FMA so do we -> agreement everywhere, kinks included
    f = _capi.FLAG_COORDS_FMA
    out = ext.ms_deform_attn_forward(*args, 64, _flags=f)
    gv, gl, ga = ext.ms_deform_attn_backward(*args, i["grad_out"], 64, _flags=f)
    torch.cuda.synchronize()
    assert rel_err(out, ref_out) < 1e-5
    assert rel_err(gv, ref_gv) < 1e-4
    assert rel_err(gl, ref_gl) < 1e-4
    assert rel_err(ga, ref_ga) < 1e-4
    # (b) default contract (mul, then sub): identical except for samples whose coordinate rounds onto
    #     the pixel lattice, where grad_sampling_loc is one-sided
    out = ext.ms_deform_attn_forward(*args, 64)
    gv, gl, ga = ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
    same = (ext.debug_corners(i["shapes"], i["starts"], i["loc"]) ==
            ext.debug_corners(i["shapes"], i["starts"], i["loc"], _flags=f)).all(-1)
    assert (~same).float().mean() < 1e-5
    assert rel_err(out, ref_out) < 1e-5
    assert rel_err(gv, ref_gv) < 1e-4
    assert rel_err(ga, ref_ga) < 1e-4
    assert rel_err(gl * same[..., None], ref_gl * same[..., None]) < 1e-4


@pytest.mark.parametrize("case", ["tiny_f32", "enc_small_f32", "pad_small_f32", "dec_small_f32"])
def test_reference_cuda_kernels_reproduce_the_golden_vectors(legacy, case):
    """Sanity of the comparator itself: the legacy kernels agree with the reference's Python checker."""
    g = load_golden(case)
    dev = "cuda:0"
    shp = torch.as_tensor(g["shape_list"], dtype=torch.long, device=dev)
    hw = shp[:, 0] * shp[:, 1]
    st = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
    out = legacy.forward(g["value"].to(dev), shp, st, g["loc"].to(dev), g["attw"].to(dev))
    assert rel_err(out.cpu(), g["out"]) < 1e-5
