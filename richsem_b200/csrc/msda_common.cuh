// msda_common.cuh — shared device helpers for the MSDeformAttn sm_100a kernels.
//
// The sample geometry below is THE index contract of the op: it restates, in fp32
// round-to-nearest with no FMA contraction, what the reference computes at
// models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:285-288 (pixel coordinates and
// range test) and :38-78 (floor, fractions, per-corner bounds).  Every kernel in this
// library (forward, backward, deterministic backward, debug probe) gets its corner
// indices from msda_sample_geom() and nowhere else.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

// Per-level table; passed by value as a __grid_constant__ kernel parameter, so it lives
// in the constant bank (c[0x0][..]) and costs no load instructions when the level index
// is known at compile time.
struct MsdaLevels {
  int H[MSDA_MAX_LEVELS];
  int W[MSDA_MAX_LEVELS];
  int start[MSDA_MAX_LEVELS];
  int coord_fma;  // MSDA_FLAG_COORDS_FMA: pixel coordinate as one fused multiply-add (see msda_pix)
};

// Fused prologue (reference: models/richsem/ops/modules/ms_deform_attn.py:98-111).
// ref_dim == 0: the kernel's `loc` / `attw` arguments are sampling locations and attention weights (the
// reference op).  ref_dim == 2 | 4: they are the RAW outputs of the module's sampling_offsets / attention_weights
// Linear layers, and the kernel itself applies the softmax over the L*P logits of a (query, head) and
//   ref_dim 2:  loc = ref[l] + offset / (W_l, H_l)                       (ms_deform_attn.py:102-105)
//   ref_dim 4:  loc = ref[l].xy + offset / P * ref[l].wh * 0.5           (ms_deform_attn.py:106-108)
// with `ref` = reference_points [N, Lq, L, ref_dim]; the backward then returns the gradients of the raw tensors.
struct MsdaFused {
  const float* ref;
  int ref_dim;
};

// Sampling location of one point from its raw offset (same operations, same order as the module's PyTorch
// expression, so the coordinates — and with them the corner indices — are bit-identical to the unfused path).
__device__ __forceinline__ float2 msda_fused_location(const MsdaFused fz, const size_t bq, const int num_levels,
                                                      const int l, const int num_point, const int H, const int W,
                                                      const float2 off) {
  const float* r = fz.ref + (bq * num_levels + l) * fz.ref_dim;
  if (fz.ref_dim == 2) return make_float2(r[0] + off.x / (float)W, r[1] + off.y / (float)H);
  return make_float2(r[0] + off.x / (float)num_point * r[2] * 0.5f, r[1] + off.y / (float)num_point * r[3] * 0.5f);
}

struct MsdaDims {
  int batch, spatial_size, num_heads, channels, num_levels, num_query, num_point;
};

// ---- pixel coordinate: loc * size - 0.5 -------------------------------------------------
// Default (the index contract): two separately rounded operations, as the source of the
// reference reads (it multiplies in scalar_t and subtracts a double literal, which for float
// is exactly a rounded multiply followed by a rounded subtract; cuh:285-286).
// fma=true: ONE fused multiply-add.  This is what nvcc actually emits for that source line
// with its default -fmad=true (the double subtraction is narrowed to float, then contracted),
// i.e. what the reference's compiled extension computes.  The two differ only when loc*size
// rounds onto the pixel lattice (k + 0.5): measured 1 sample in 11.4 M for jittered encoder
// locations, but every lattice point of an un-jittered initialisation.
__device__ __forceinline__ float msda_pix(float loc, int size, bool fma) {
  return fma ? __fmaf_rn(loc, (float)size, -0.5f) : __fsub_rn(__fmul_rn(loc, (float)size), 0.5f);
}
__device__ __forceinline__ double msda_pix(double loc, int size, bool fma) {
  return fma ? __fma_rn(loc, (double)size, -0.5) : __dsub_rn(__dmul_rn(loc, (double)size), 0.5);
}

// Geometry of one sampling point.
//   tok[k]  token index (into the spatial_size axis) of corner k in the order
//           (h0,w0) (h0,w1) (h1,w0) (h1,w1); -1 when the corner contributes nothing.
//   lh, lw  fractional parts (0 when the whole sample is skipped).
// Returns true when the sample passes the range test of cuh:288.
template <typename T>
__device__ __forceinline__ bool msda_sample_geom_hw(T x, T y, int H, int W, int start, int (&tok)[4],
                                                    T& lh, T& lw, int& h0, int& w0, bool fma = false) {
  const T w_im = msda_pix(x, W, fma);
  const T h_im = msda_pix(y, H, fma);
  tok[0] = tok[1] = tok[2] = tok[3] = -1;
  lh = T(0);
  lw = T(0);
  h0 = w0 = 0;
  // NaN coordinates fail every comparison and are skipped, as in the reference.
  if (!(h_im > T(-1) && w_im > T(-1) && h_im < T(H) && w_im < T(W))) return false;
  const T hf = floor(h_im);
  const T wf = floor(w_im);
  h0 = (int)hf;
  w0 = (int)wf;
  lh = h_im - hf;
  lw = w_im - wf;
  const bool h0ok = h0 >= 0, w0ok = w0 >= 0;
  const bool h1ok = h0 + 1 <= H - 1, w1ok = w0 + 1 <= W - 1;
  const int base = start + h0 * W + w0;
  if (h0ok && w0ok) tok[0] = base;
  if (h0ok && w1ok) tok[1] = base + 1;
  if (h1ok && w0ok) tok[2] = base + W;
  if (h1ok && w1ok) tok[3] = base + W + 1;
  return true;
}

template <typename T>
__device__ __forceinline__ bool msda_sample_geom(T x, T y, int H, int W, int start, int (&tok)[4],
                                                 T& lh, T& lw, bool fma = false) {
  int h0, w0;
  return msda_sample_geom_hw(x, y, H, W, start, tok, lh, lw, h0, w0, fma);
}

// ---- packed fp32 math (Blackwell FFMA2: two fp32 FMAs per issued instruction) --------------
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// ---- memory helpers ---------------------------------------------------------------------
// Streaming operands (sampling_loc, attn_weight, grad_out, outputs) are touched once:
// keep them out of L1 so the gathered `value` rows own it.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_stream_f2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f1(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_f2(float* p, float2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
// 16-byte fp32 vector reduction: one REDG.E.ADD.F32x4 per lane (sm_90+).  No "memory" clobber: the kernels never
// read grad_value, and with the clobber the compiler could not hoist the next corner's row load above the
// previous corner's reduction (a chain of dependent L2 round trips, profiles/r2_bwd_window_phase_budget.md).
// volatile keeps the reductions in program order relative to griddepcontrol.wait.
__device__ __forceinline__ void red_add_f4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}
// Two 16-byte reductions into baseA[off .. off+3] and baseB[off .. off+3] (off in elements), skipped when off < 0.
// Branch-free: one predicate, one IMAD.WIDE per address (the compiler's own code for the same thing is a branch
// with reconvergence barriers and a four-instruction 64-bit address computation per reduction).
__device__ __forceinline__ void red_add_2xf4_if(float* baseA, float* baseB, int off, float a0, float a1, float a2,
                                                float a3, float b0, float b1, float b2, float b3) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .u64 pa, pb;\n"
      " setp.ge.s32 p, %2, 0;\n"
      " mad.wide.s32 pa, %2, 4, %0;\n"
      " mad.wide.s32 pb, %2, 4, %1;\n"
      " @p red.global.add.v4.f32 [pa], {%3,%4,%5,%6};\n"
      " @p red.global.add.v4.f32 [pb], {%7,%8,%9,%10};\n}"
      ::"l"(baseA), "l"(baseB), "r"(off), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(b0), "f"(b1), "f"(b2), "f"(b3));
}

