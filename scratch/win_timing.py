"""Phase timing of the window kernels (needs a library built with MSDA_NVCC_EXTRA=-DMSDA_WIN_TIMING)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from richsem_b200 import _capi, synthetic as syn, MultiScaleDeformableAttention as ext
lib = _capi.lib
lib.msda_debug_win_timing.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
shapes = syn.level_shapes(800, 1333)
i = syn.make_inputs("E", 2, shapes, "cuda:0", seed=1)
args = (i["value"], i["shapes"], i["starts"], i["loc"], i["attw"])
buf = (ctypes.c_uint64 * 16)()
for _ in range(3):
    ext.ms_deform_attn_forward(*args, 64); ext.ms_deform_attn_backward(*args, i["grad_out"], 64)
lib.msda_debug_win_timing(buf)
n = 5
for _ in range(n):
    ext.ms_deform_attn_forward(*args, 64)
lib.msda_debug_win_timing(buf); t = list(buf)
nb = max(t[7], 1)
print("fwd blocks", nb // n, "cycles/block: decode+bbox %.0f | alloc+stage+records %.0f | window wait %.0f | gather %.0f" % tuple(x / nb for x in t[0:4]))
print("  fwd gather: blocks with a direct level %d of %d, their gather %.0f cycles, fully windowed blocks %.0f" % (t[5] // n, nb // n, t[4] / max(t[5], 1), (t[3] - t[4]) / max(nb - t[5], 1)))
for _ in range(n):
    ext.ms_deform_attn_backward(*args, i["grad_out"], 64, _flags=_capi.FLAG_BWD_WS)
lib.msda_debug_win_timing(buf); t = list(buf)
nb = max(t[15], 1)
print("bwd warp-specialised, tiles", nb // n, "cycles/tile: produce %.0f | consumer waits %.0f | consume %.0f | producer waits %.0f" % tuple(x / nb for x in t[8:12]))
