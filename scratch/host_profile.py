"""cProfile of the Python launch path (GPU box): where do the ~40 us per call go?"""
import cProfile, pstats, sys, io
sys.path.insert(0, ".")
import torch
from richsem_b200 import MultiScaleDeformableAttention as ext, synthetic as syn
i = syn.make_inputs("Dn", 2, syn.level_shapes(800, 1333), "cuda:0", lq=8)
args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
for _ in range(300):
    ext.ms_deform_attn_forward(*args, 64); ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    ext.ms_deform_attn_forward(*args, 64)
    ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
