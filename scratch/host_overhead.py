"""Host time per call of the drop-in API on the decoder workload: the Python loop that enqueues one step (6 forwards +
6 backwards) is timed with the wall clock and compared with the step's device time."""
import sys, time, torch
sys.path.insert(0, '.')
from richsem_b200 import MultiScaleDeformableAttention as ext, synthetic as syn
shapes = syn.level_shapes(800, 1333)
sets = [syn.make_inputs("Dn", 2, shapes, "cuda:0", seed=k, lq=1100, dtype=torch.bfloat16) for k in range(6)]
shp, st = sets[0]["shapes"], sets[0]["starts"]
def step():
    for s in sets:
        s["out"] = ext.ms_deform_attn_forward(s["value"], shp, st, s["loc"], s["attw"], 64)
    for s in reversed(sets):
        s["g"] = ext.ms_deform_attn_backward(s["value"], shp, st, s["loc"], s["attw"], s["grad_out"], 64)
for _ in range(20): step()
torch.cuda.synchronize()
# host time alone: enqueue with an idle-enough GPU queue (sync first), many steps, no sync inside
N = 200
t0 = time.perf_counter()
for _ in range(N): step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"enqueue {1e6*t_host/N/12:.1f} us per call; {1e3*t_all/N:.4f} ms per step wall incl. drain")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
