"""Builds richsem_b200/lib/libmsda_b200.so with nvcc for sm_100a (no GPU needed to compile).

Every csrc/*.cu is one translation unit; they are compiled in parallel into lib/obj/*.o (only the
stale ones) and linked into the shared library.  ptxas -v output of all units is kept in
lib/ptxas_info.txt (registers / spills / shared memory per kernel)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libmsda_b200.so"
OBJ = PKG / "lib" / "obj"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-Xptxas=-v",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(CSRC.glob("*.cu")), sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                                             [PKG.parent / "include" / "msda_b200.h"])


def _deps(path: Path, seen=None):
    """Transitive closure of the quoted #includes of `path` (inside the repo)."""
    import re

    seen = set() if seen is None else seen
    if path in seen or not path.exists():
        return seen
    seen.add(path)
    for inc in re.findall(r'^\s*#include\s+"([^"]+)"', path.read_text(), flags=re.M):
        _deps((path.parent / inc).resolve(), seen)
    return seen


def _compile(nvcc, cu, obj, extra):
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", str(obj), str(cu)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    obj.with_suffix(".ptxas.txt").write_text(res.stderr)
    return res.stderr


def build_library(force: bool = False, verbose: bool = False) -> Path:
    cus, deps = sources()
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("MSDA_NVCC_EXTRA", "").split()
    OBJ.mkdir(parents=True, exist_ok=True)
    stale = []
    for cu in cus:
        obj = OBJ / (cu.stem + ".o")
        newest = max(p.stat().st_mtime for p in _deps(cu.resolve()))
        if force or extra or not obj.exists() or obj.stat().st_mtime < newest:
            stale.append((cu, obj))
    for old in OBJ.glob("*.o"):  # a removed source must not leave its object behind
        if old.stem not in {cu.stem for cu in cus}:
            old.unlink()
            stale = stale or []
            force = True
    if not stale and not force and LIB.exists() and all(LIB.stat().st_mtime >= (OBJ / (cu.stem + ".o")).stat().st_mtime
                                                        for cu in cus):
        return LIB
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(stale)))) as ex:
        list(ex.map(lambda co: _compile(nvcc, co[0], co[1], extra), stale))
    objs = [str(OBJ / (cu.stem + ".o")) for cu in cus]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed: " + " ".join(cmd))
    info = "".join((OBJ / (cu.stem + ".ptxas.txt")).read_text() for cu in cus)
    (LIB.parent / "ptxas_info.txt").write_text(info)
    if verbose:
        print(info)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
