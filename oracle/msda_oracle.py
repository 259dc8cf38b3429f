"""CPU oracle for RichSem's MSDeformAttn.  TEST INFRASTRUCTURE ONLY — never on the product path.

Two independent restatements of the reference algorithm:

* ``core_pytorch`` — the grid_sample formulation of the reference's own checker
  ``ms_deform_attn_core_pytorch`` (/root/reference/models/richsem/ops/functions/ms_deform_attn_func.py:41-61),
  re-derived here because /root/reference does not exist on the GPU box.  Differentiable, so autograd
  through it is also the backward oracle.  PINNED: tests/test_oracle.py asserts it is bit-identical to
  the reference function (loaded by path when /root/reference is present) and reproduces the golden
  vectors in tests/golden/ that were generated from the reference function.
* ``COracle`` — ctypes front end of oracle/msda_oracle.c, a scalar C restatement of the reference CUDA
  kernels' arithmetic (ms_deform_im2col_cuda.cuh:33-159, 237-299) that additionally exposes the bilinear
  corner indices (the bit-exact contract).  PINNED against the same golden vectors (tolerance for
  values, exact for indices derived from them).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "libmsda_oracle.so"


# --------------------------------------------------------------------------------------
# grid_sample formulation  (func.py:41-61)
# --------------------------------------------------------------------------------------
def core_pytorch(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D); spatial_shapes iterable of (H,W); sampling_locations (N,Lq,M,L,P,2) in
    [0,1] as (x,y); attention_weights (N,Lq,M,L,P)  ->  (N, Lq, M*D)."""
    n, _, m, d = value.shape
    lq, n_lvl, n_pts = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    sizes = [(int(h), int(w)) for h, w in spatial_shapes]
    per_level = value.split([h * w for h, w in sizes], dim=1)
    grids = 2 * sampling_locations - 1  # grid_sample's [-1,1] convention, align_corners=False (func.py:47)
    sampled = []
    for lvl, (h, w) in enumerate(sizes):
        # (N, HW, M, D) -> (N*M, D, H, W): one image of D channels per (sample, head)   (func.py:51)
        img = per_level[lvl].flatten(2).transpose(1, 2).reshape(n * m, d, h, w)
        # (N, Lq, M, P, 2) -> (N*M, Lq, P, 2)                                            (func.py:53)
        grid = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
    # weights (N, Lq, M, L, P) -> (N*M, 1, Lq, L*P)                                      (func.py:59)
    wts = attention_weights.transpose(1, 2).reshape(n * m, 1, lq, n_lvl * n_pts)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * wts).sum(-1).view(n, m * d, lq)  # (func.py:60)
    return out.transpose(1, 2).contiguous()


def core_pytorch_fwd_bwd(value, spatial_shapes, loc, attw, grad_out):
    """Forward + autograd backward of ``core_pytorch``; returns (out, grad_value, grad_loc, grad_attw)."""
    v = value.detach().clone().requires_grad_(True)
    l = loc.detach().clone().requires_grad_(True)
    a = attw.detach().clone().requires_grad_(True)
    out = core_pytorch(v, spatial_shapes, l, a)
    out.backward(grad_out)
    return out.detach(), v.grad, l.grad, a.grad


def load_reference_core(reference_root="/root/reference"):
    """The reference's own ms_deform_attn_core_pytorch, loaded by path with its CUDA extension import
    stubbed.  Returns None when the reference tree is absent (e.g. on the GPU box)."""
    import importlib.util
    import sys
    import types

    path = Path(reference_root) / "models/richsem/ops/functions/ms_deform_attn_func.py"
    if not path.exists():
        return None
    stub_name = "MultiScaleDeformableAttention"
    had = sys.modules.get(stub_name)
    sys.modules[stub_name] = types.ModuleType(stub_name)
    try:
        spec = importlib.util.spec_from_file_location("_richsem_ref_msda_func", str(path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if had is None:
            sys.modules.pop(stub_name, None)
        else:
            sys.modules[stub_name] = had
    return mod.ms_deform_attn_core_pytorch


# --------------------------------------------------------------------------------------
# C restatement
# --------------------------------------------------------------------------------------
def build(force=False):
    """Compile oracle/msda_oracle.c with gcc (idempotent)."""
    src = _HERE / "msda_oracle.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s", "_build/libmsda_oracle.so"],
                       check=True, stdout=subprocess.DEVNULL)
    return _LIB


class COracle:
    """ctypes wrapper over libmsda_oracle.so; takes / returns CPU torch tensors (fp32 or fp64)."""

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(str(_LIB))
        for f in ("msda_oracle_forward_f32", "msda_oracle_forward_f64", "msda_oracle_backward_f32",
                  "msda_oracle_backward_f64", "msda_oracle_corners_f32", "msda_oracle_corners_f32_fma"):
            getattr(self.lib, f).restype = None

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def _meta(spatial_shapes, level_start_index=None):
        shp = torch.as_tensor(np.asarray([[int(h), int(w)] for h, w in spatial_shapes]), dtype=torch.int64)
        if level_start_index is None:
            hw = shp[:, 0] * shp[:, 1]
            st = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
        else:
            st = torch.as_tensor(level_start_index, dtype=torch.int64).cpu().contiguous()
        return shp.contiguous(), st.contiguous()

    def forward(self, value, spatial_shapes, loc, attw, level_start_index=None):
        value, loc, attw = value.contiguous(), loc.contiguous(), attw.contiguous()
        n, s, m, d = value.shape
        lq, nl, npt = loc.shape[1], loc.shape[3], loc.shape[4]
        shp, st = self._meta(spatial_shapes, level_start_index)
        out = torch.empty(n, lq, m * d, dtype=value.dtype)
        fn = self.lib.msda_oracle_forward_f32 if value.dtype == torch.float32 else self.lib.msda_oracle_forward_f64
        fn(self._p(value), self._p(shp), self._p(st), self._p(loc), self._p(attw),
           n, s, m, d, nl, lq, npt, self._p(out))
        return out

    def backward(self, grad_out, value, spatial_shapes, loc, attw, level_start_index=None):
        grad_out, value, loc, attw = grad_out.contiguous(), value.contiguous(), loc.contiguous(), attw.contiguous()
        n, s, m, d = value.shape
        lq, nl, npt = loc.shape[1], loc.shape[3], loc.shape[4]
        shp, st = self._meta(spatial_shapes, level_start_index)
        gv, gl, ga = torch.zeros_like(value), torch.empty_like(loc), torch.empty_like(attw)
        fn = self.lib.msda_oracle_backward_f32 if value.dtype == torch.float32 else self.lib.msda_oracle_backward_f64
        fn(self._p(grad_out), self._p(value), self._p(shp), self._p(st), self._p(loc), self._p(attw),
           n, s, m, d, nl, lq, npt, self._p(gv), self._p(gl), self._p(ga))
        return gv, gl, ga

    def corners(self, spatial_shapes, loc, level_start_index=None, fma=False):
        """int32 (N,Lq,M,L,P,4): token index of each bilinear corner, -1 where it contributes nothing.
        fma=True: pixel coordinate as one fused multiply-add (the compiled reference kernel's arithmetic)."""
        loc = loc.contiguous().float()
        n, lq, m, nl, npt, _ = loc.shape
        shp, st = self._meta(spatial_shapes, level_start_index)
        out = torch.empty(n, lq, m, nl, npt, 4, dtype=torch.int32)
        fn = self.lib.msda_oracle_corners_f32_fma if fma else self.lib.msda_oracle_corners_f32
        fn(self._p(shp), self._p(st), self._p(loc), ctypes.c_int64(n * lq * m), nl, npt, self._p(out))
        return out


def corners_numpy(spatial_shapes, loc, level_start_index=None, fma=False):
    """Pure-numpy cross-check of COracle.corners (fp32 mul, then fp32 sub, then floor); with fma=True the
    coordinate is the exactly computed x*W - 0.5 (float64 holds it exactly) rounded once to fp32."""
    loc = np.asarray(loc, dtype=np.float32)
    n, lq, m, nl, npt, _ = loc.shape
    out = np.full((n, lq, m, nl, npt, 4), -1, dtype=np.int32)
    start = 0
    for l, (h, w) in enumerate(spatial_shapes):
        h, w = int(h), int(w)
        st = start if level_start_index is None else int(level_start_index[l])
        start += h * w
        if fma:
            wim = (loc[:, :, :, l, :, 0].astype(np.float64) * w - 0.5).astype(np.float32)
            him = (loc[:, :, :, l, :, 1].astype(np.float64) * h - 0.5).astype(np.float32)
        else:
            wim = (loc[:, :, :, l, :, 0] * np.float32(w)).astype(np.float32) - np.float32(0.5)
            him = (loc[:, :, :, l, :, 1] * np.float32(h)).astype(np.float32) - np.float32(0.5)
        ok = (him > -1) & (wim > -1) & (him < h) & (wim < w)
        h0 = np.floor(him).astype(np.int64)
        w0 = np.floor(wim).astype(np.int64)
        for k, (dh, dw) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            hh, ww = h0 + dh, w0 + dw
            good = ok & (hh >= 0) & (hh <= h - 1) & (ww >= 0) & (ww <= w - 1)
            out[:, :, :, l, :, k] = np.where(good, st + hh * w + ww, -1)
    return out


# --------------------------------------------------------------------------------------
# The reference's own CUDA kernels (oracle/_ref/libmsda_legacy.so, built by oracle/build_ref.py
# from /root/reference in the build container).  GPU comparator: second parity check and
# "legacy kernel recompiled for sm_100a" timing.  Never the product path.
# --------------------------------------------------------------------------------------
_LEGACY = _HERE / "_ref" / "libmsda_legacy.so"


class LegacyCuda:
    """ctypes front end of the reference launchers (cuh:923-954, 956-1327); CUDA fp32 tensors."""

    @staticmethod
    def available():
        return _LEGACY.exists()

    def __init__(self):
        self.lib = ctypes.CDLL(str(_LEGACY))
        self.lib.legacy_forward_f32.restype = None
        self.lib.legacy_backward_f32.restype = None

    @staticmethod
    def _args(value, shapes, starts, loc, attw):
        n, s, m, d = value.shape
        lq, nl, npt = loc.shape[1], loc.shape[3], loc.shape[4]
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        return stream, p, (n, s, m, d, nl, lq, npt)

    def forward(self, value, shapes, starts, loc, attw):
        stream, p, dims = self._args(value, shapes, starts, loc, attw)
        n, s, m, d, nl, lq, npt = dims
        out = torch.zeros(n, lq, m * d, dtype=value.dtype, device=value.device)  # reference: at::zeros (cu:54)
        self.lib.legacy_forward_f32(stream, p(value), p(shapes), p(starts), p(loc), p(attw), *dims, p(out))
        return out

    def backward(self, value, shapes, starts, loc, attw, grad_out):
        stream, p, dims = self._args(value, shapes, starts, loc, attw)
        gv, gl, ga = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(attw)  # cu:121-123
        self.lib.legacy_backward_f32(stream, p(grad_out), p(value), p(shapes), p(starts), p(loc), p(attw), *dims,
                                     p(gv), p(gl), p(ga))
        return gv, gl, ga
