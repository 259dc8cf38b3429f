"""Times the REFERENCE's CUDA kernels (recompiled for sm_100a, oracle/_ref) next to ours on the same
inputs and prints a markdown table (-> profiles/).  Comparator script, not a pytest module.

    python tests/perf_legacy_compare.py > gpurun_out/legacy_vs_b200.md
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.msda_oracle import LegacyCuda  # noqa: E402
from richsem_b200 import MultiScaleDeformableAttention as ext, synthetic as syn  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    legacy = LegacyCuda()
    rows = []
    for name, hw, kind, lq in (("encoder 800x1333 bs=2", (800, 1333), "E", None),
                               ("encoder 1600x2000 bs=2", (1600, 2000), "E", None),
                               ("decoder 800x1333 bs=2 Lq=1100", (800, 1333), "Dn", 1100)):
        shapes = syn.level_shapes(*hw)
        sets = [syn.make_inputs(kind, 2, shapes, "cuda:0", seed=10 + k, lq=lq) for k in range(4)]  # > L2 in total
        k = [0]

        def nxt():
            k[0] = (k[0] + 1) % len(sets)
            s = sets[k[0]]
            return (s["value"], s["shapes"], s["starts"], s["loc"], s["attw"]), s["grad_out"]

        def ours_f():
            a, _ = nxt(); ext.ms_deform_attn_forward(*a, 64)

        def ours_b():
            a, g = nxt(); ext.ms_deform_attn_backward(*a, g, 64)

        def leg_f():
            a, _ = nxt(); legacy.forward(*a)

        def leg_b():
            a, g = nxt(); legacy.backward(*a, g)

        rows.append((name, timeit(leg_f), timeit(ours_f), timeit(leg_b), timeit(ours_b)))
    print("| shape (fp32, M=8, D=32, L=4, P=4) | legacy fwd ms | ours fwd ms | x | legacy bwd ms | ours bwd ms | x | fwd+bwd x |")
    print("|---|---|---|---|---|---|---|---|")
    for n, lf, of, lb, ob in rows:
        print(f"| {n} | {lf:.3f} | {of:.3f} | {lf / of:.2f} | {lb:.3f} | {ob:.3f} | {lb / ob:.2f} | {(lf + lb) / (of + ob):.2f} |")
    print("\nlegacy = the reference's ms_deform_im2col_cuda.cuh kernels compiled unmodified for sm_100a "
          "(oracle/build_ref.py), including the zero-fills its host wrapper performs "
          "(ms_deform_attn_cuda.cu:54,121-123); ours = libmsda_b200 through the drop-in Python API. "
          "CUDA events, 20 iterations after 5 warm-ups, 4 rotating input sets.")


if __name__ == "__main__":
    main()
