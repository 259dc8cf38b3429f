"""Debug helper (not a test): this library, the C oracle and the reference CUDA kernels (oracle/_ref) side by side on one
encoder layer.  Lives under tests/ because only tests may use oracle/.   python tests/dbg_legacy_compare.py"""
import sys; sys.path.insert(0, ".")
import torch
from oracle.msda_oracle import LegacyCuda, COracle
from richsem_b200 import MultiScaleDeformableAttention as ext, synthetic as syn
shapes = syn.level_shapes(800, 1333)
i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=77)
args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
gv, gl, ga = ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
leg = LegacyCuda()
rgv, rgl, rga = leg.backward(*args, i["grad_out"])
torch.cuda.synchronize()
d = (gl - rgl).abs()
print("max|gl|", rgl.abs().max().item(), "max diff", d.max().item(), "n > 1e-3*max:", (d > 1e-3 * rgl.abs().max()).sum().item(), "of", d.numel())
idx = torch.nonzero(d > 1e-3 * rgl.abs().max())[:8]
co = COracle()
for t in idx.tolist():
    b, q, m, l, p, xy = t
    loc = i["loc"][b, q, m, l, p]
    H, W = shapes[l]
    wim = (loc[0] * W - 0.5).item(); him = (loc[1] * H - 0.5).item()
    print(t, "loc", loc.tolist(), "w_im", wim, "h_im", him, "ours", gl[b, q, m, l, p].tolist(), "legacy", rgl[b, q, m, l, p].tolist(), "ga ours/legacy", ga[b,q,m,l,p].item(), rga[b,q,m,l,p].item())
# C oracle on image 0 only
v, loc, w, go = (i[k][:1].cpu() for k in ("value", "loc", "attw", "grad_out"))
cgv, cgl, cga = co.backward(go, v, shapes, loc, w)
print("C oracle vs ours (img0):", ((cgl - gl[:1].cpu()).abs().max() / cgl.abs().max()).item(), " C oracle vs legacy:", ((cgl - rgl[:1].cpu()).abs().max() / cgl.abs().max()).item())
