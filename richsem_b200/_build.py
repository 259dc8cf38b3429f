"""Builds richsem_b200/lib/libmsda_b200.so with nvcc for sm_100a (no GPU needed to compile)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libmsda_b200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-Xptxas=-v",
    "-shared", "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(CSRC.glob("*.cu")), sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                                             [PKG.parent / "include" / "msda_b200.h"])


def build_library(force: bool = False, verbose: bool = False) -> Path:
    cus, deps = sources()
    newest = max(p.stat().st_mtime for p in cus + deps)
    if not force and LIB.exists() and LIB.stat().st_mtime >= newest:
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB), *map(str, cus)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    (LIB.parent / "ptxas_info.txt").write_text(res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
