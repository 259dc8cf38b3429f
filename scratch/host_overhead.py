"""Host-side cost per call of the drop-in API (GPU box): tiny problem so the GPU is never the limit."""
import sys, time
sys.path.insert(0, ".")
import torch
from richsem_b200 import MultiScaleDeformableAttention as ext, synthetic as syn
i = syn.make_inputs("Dn", 2, syn.level_shapes(800, 1333), "cuda:0", lq=8)
args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
for name, fn in (("forward", lambda: ext.ms_deform_attn_forward(*args, 64)),
                 ("backward", lambda: ext.ms_deform_attn_backward(*args, i["grad_out"], 64))):
    for _ in range(200): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2000): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name}: {(t1 - t0) / 2000 * 1e6:.1f} us host time per call")
