"""Per-warp phase timing of the window backward (build the variant with scratch/build_variant.sh timing msda_launch_win
"-DMSDA_WIN_TIMING" and run with MSDA_B200_LIB pointing at it)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from richsem_b200 import _capi, MultiScaleDeformableAttention as ext, synthetic as syn
shapes = syn.level_shapes(800, 1333)
i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=1)
args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
buf = (ctypes.c_ulonglong * 32)()
for _ in range(3):
    ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
_capi.lib.msda_debug_win_timing(buf)
ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
_capi.lib.msda_debug_win_timing(buf)
nb = buf[17]
print("blocks", nb, "front end cycles/block", buf[16] / nb)
print("sorted pass cycles/block per warp:", [round(buf[w] / nb) for w in range(8)])
print("direct pass cycles/block per warp:", [round(buf[8 + w] / nb) for w in range(8)])
names = ["loads issued + init + barrier", "decode (waits for loads) + bbox + barrier", "alloc + staging issue + records + counts + barrier",
         "scan (2 barriers)", "placement", "cp.async wait (thread 0)", "final barrier"]
for k, nm in enumerate(names):
    print(f"  front end: {nm}: {buf[20 + k] / nb:.0f} cycles/block")
