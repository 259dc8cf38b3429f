"""Small forward+backward through every D=32 kernel family (for compute-sanitizer)."""
import sys, torch
sys.path.insert(0, '.')
from richsem_b200 import _capi, synthetic as syn, MultiScaleDeformableAttention as ext
shapes = [(20, 31), (10, 16), (5, 8), (3, 4)]
for kind, lq in (("E", None), ("U", 300), ("Dn", 200)):
    for dt in (torch.float32, torch.bfloat16):
        i = syn.make_inputs(kind, 2, shapes, "cuda:0", seed=3, lq=lq, dtype=dt)
        if kind == "U":
            i["loc"] = (i["loc"] * 1.5 - 0.25).contiguous()
        a = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
        for kern in (_capi.KERNEL_AUTO, _capi.KERNEL_SPLIT, _capi.KERNEL_TILED, _capi.KERNEL_WINDOW):
            for fl in (0, _capi.FLAG_DETERMINISTIC):
                o = ext.ms_deform_attn_forward(*a, 64, _kernel=kern)
                g = ext.ms_deform_attn_backward(*a, i["grad_out"], 64, _flags=fl, _kernel=kern)
torch.cuda.synchronize()
print("ok")
