"""Builds oracle/_ref/libmsda_legacy.so: the REFERENCE's CUDA kernels for this path, compiled for
sm_100a from the sources where they lie under /root/reference (never copied), through
oracle/ref_legacy_shim.cu.  Only possible where the reference tree exists (the build container);
the built .so is git-ignored but travels to the GPU box.  It is a comparator for tests / bench
(GPU parity of the legacy kernel vs ours, and "legacy kernel on B200" timing), never product code.

The reference's own build (models/richsem/ops/setup.py) is not used: it refuses to run without a
visible GPU (:48-49) and its host wrappers no longer compile against torch 2.11
(ms_deform_attn_cuda.cu:64,134).  The kernel header itself needs nothing from torch: the three
torch includes at its top are satisfied by empty stubs in oracle/ref_stubs/.
"""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/models/richsem/ops/src/cuda")
OUT = HERE / "_ref" / "libmsda_legacy.so"


def main():
    if not (REF / "ms_deform_im2col_cuda.cuh").exists():
        print("[build_ref] reference tree absent; keeping any prebuilt oracle/_ref")
        return 0
    src = HERE / "ref_legacy_shim.cu"
    if OUT.exists() and OUT.stat().st_mtime >= max(src.stat().st_mtime, (REF / "ms_deform_im2col_cuda.cuh").stat().st_mtime):
        print(f"[build_ref] up to date: {OUT}")
        return 0
    OUT.parent.mkdir(exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
           "-w", "-I", str(HERE / "ref_stubs"), "-I", str(REF), "-o", str(OUT), str(src)]
    subprocess.run(cmd, check=True)
    print(f"[build_ref] built {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
