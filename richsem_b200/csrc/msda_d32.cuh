// msda_d32.cuh — tuned fp32 kernels for head_dim (channels) == 32.
//
// Work decomposition (both directions)
//   grid  = (num_heads * tiles, batch); one thread block = ONE head x a tile of
//           kTileQ queries taken from `order` (patch-tiled for encoder self-attention,
//           natural order otherwise).  Heads sample in different directions, so a block
//           that sticks to one head keeps the rows it gathers resident in L1.
//   warp  = 32/G "lane groups"; one lane group = one (query, head) pair.  A group of G
//           lanes covers the 32 channels of a value row with one vector load per lane
//           (G=8: LDG.128, G=4: LDG.256), so a row is always one full 128-byte line.
//   stage 1 each lane decodes two sampling points per 2G points of its (query, head):
//           coalesced float4 of sampling_loc + float2 of attn_weight, bit-exact geometry
//           (msda_sample_geom), and publishes a 32-byte record per point to shared memory.
//   stage 2 all lanes of the group walk the L*P records (broadcast LDS.128), issue the
//           four predicated row gathers and blend.
// Algorithm restated from models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299
// (forward), :87-159 and :301-403 (backward); nothing is shared with that code's
// thread mapping (one thread per output channel, one-warp blocks, serial reductions).
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTileQ = 64;  // queries per thread block

template <int G, int LP>
struct D32Cfg {
  static constexpr int C = 32 / G;              // channels per lane
  static constexpr int GPW = 32 / G;            // (query, head) pairs per warp
  static constexpr int QPP = kWarps * GPW;      // queries per pass of the block
  static constexpr int PASSES = kTileQ / QPP;
  static constexpr int KP = (LP + 2 * G - 1) / (2 * G);  // point pairs decoded per lane
  static constexpr int REC_STRIDE = 2 * LP + 1;  // int4 units per lane group (+1: bank skew)
  static constexpr int SMEM_BYTES = kWarps * GPW * REC_STRIDE * 16;
  static_assert(LP % 2 == 0, "L*P must be even");
  static_assert(kTileQ % QPP == 0, "tile must be a whole number of passes");
};

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int G, int kL, int kP>
__global__ void __launch_bounds__(kThreads)
msda_fwd_d32_kernel(const float* __restrict__ value, const float* __restrict__ loc,
                    const float* __restrict__ attw, float* __restrict__ out,
                    const int* __restrict__ order, const int order_len,
                    const __grid_constant__ MsdaLevels lv, const int S, const int M, const int Lq) {
  constexpr int LP = kL * kP;
  using Cfg = D32Cfg<G, LP>;
  constexpr int C = Cfg::C;
  extern __shared__ int4 smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, j = lane % G;
  const int m = blockIdx.x % M;
  const int tile = blockIdx.x / M;
  const int b = blockIdx.y;

  int4* rec_off = smem + (warp * Cfg::GPW + g) * Cfg::REC_STRIDE;
  float4* rec_w = reinterpret_cast<float4*>(rec_off + LP);
  const float* value_b = value + (size_t)b * S * M * 32 + j * C;

#pragma unroll 1
  for (int pass = 0; pass < Cfg::PASSES; ++pass) {
    const int slot = tile * kTileQ + pass * Cfg::QPP + warp * Cfg::GPW + g;
    int q = -1;
    if (slot < order_len) q = order ? order[slot] : slot;
    const bool active = q >= 0;
    const size_t qm = ((size_t)b * Lq + (active ? q : 0)) * M + m;

    // ---- stage 1: decode this lane's points -------------------------------------------
    if (active) {
#pragma unroll
      for (int k = 0; k < Cfg::KP; ++k) {
        const int p0 = 2 * j + 2 * G * k;
        if (p0 < LP) {
          const float4 xy = ld_stream_f4(loc + qm * (LP * 2) + p0 * 2);
          const float2 aw = ld_stream_f2(attw + qm * LP + p0);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int p = p0 + i;
            const int l = p / kP;
            const float x = i ? xy.z : xy.x, y = i ? xy.w : xy.y, a = i ? aw.y : aw.x;
            int tok[4];
            float lh, lw;
            msda_sample_geom(x, y, lv.H[l], lv.W[l], lv.start[l], tok, lh, lw);
            const float hh = 1.f - lh, hw = 1.f - lw;
            int4 o;
            o.x = tok[0] >= 0 ? (tok[0] * M + m) * 32 : -1;
            o.y = tok[1] >= 0 ? (tok[1] * M + m) * 32 : -1;
            o.z = tok[2] >= 0 ? (tok[2] * M + m) * 32 : -1;
            o.w = tok[3] >= 0 ? (tok[3] * M + m) * 32 : -1;
            rec_off[p] = o;
            rec_w[p] = make_float4(a * (hh * hw), a * (hh * lw), a * (lh * hw), a * (lh * lw));
          }
        }
      }
    }
    __syncwarp();

    // ---- stage 2: gather + blend -------------------------------------------------------
    if (active) {
      float acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
      for (int p = 0; p < LP; ++p) {
        const int4 o = rec_off[p];
        const float4 w = rec_w[p];
        RowFrag<C> r0, r1, r2, r3;
        row_load_or_zero(r0, value_b, o.x);
        row_load_or_zero(r1, value_b, o.y);
        row_load_or_zero(r2, value_b, o.z);
        row_load_or_zero(r3, value_b, o.w);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          acc[c] = fmaf(w.x, r0.v[c], acc[c]);
          acc[c] = fmaf(w.y, r1.v[c], acc[c]);
          acc[c] = fmaf(w.z, r2.v[c], acc[c]);
          acc[c] = fmaf(w.w, r3.v[c], acc[c]);
        }
      }
      float* o_ptr = out + qm * 32 + j * C;
#pragma unroll
      for (int c = 0; c < C; c += 4)
        st_stream_f4(o_ptr + c, make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// backward, atomic grad_value (REDG.E.ADD.F32x4); kScatter=false leaves grad_value alone
// (deterministic mode computes it separately, msda_det.cuh)
// ------------------------------------------------------------------------------------------
template <int G, int kL, int kP, bool kScatter>
__global__ void __launch_bounds__(kThreads)
msda_bwd_d32_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                    const float* __restrict__ loc, const float* __restrict__ attw,
                    float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_attw, const int* __restrict__ order,
                    const int order_len, const __grid_constant__ MsdaLevels lv, const int S,
                    const int M, const int Lq) {
  constexpr int LP = kL * kP;
  using Cfg = D32Cfg<G, LP>;
  constexpr int C = Cfg::C;
  extern __shared__ int4 smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, j = lane % G;
  const int m = blockIdx.x % M;
  const int tile = blockIdx.x / M;
  const int b = blockIdx.y;

  int4* rec_off = smem + (warp * Cfg::GPW + g) * Cfg::REC_STRIDE;
  float4* rec_f = reinterpret_cast<float4*>(rec_off + LP);  // (lh, lw, a, -)
  const float* value_b = value + (size_t)b * S * M * 32 + j * C;
  float* gvalue_b = grad_value + (size_t)b * S * M * 32 + j * C;

#pragma unroll 1
  for (int pass = 0; pass < Cfg::PASSES; ++pass) {
    const int slot = tile * kTileQ + pass * Cfg::QPP + warp * Cfg::GPW + g;
    int q = -1;
    if (slot < order_len) q = order ? order[slot] : slot;
    const bool active = q >= 0;
    const size_t qm = ((size_t)b * Lq + (active ? q : 0)) * M + m;

    if (active) {
#pragma unroll
      for (int k = 0; k < Cfg::KP; ++k) {
        const int p0 = 2 * j + 2 * G * k;
        if (p0 < LP) {
          const float4 xy = ld_stream_f4(loc + qm * (LP * 2) + p0 * 2);
          const float2 aw = ld_stream_f2(attw + qm * LP + p0);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int p = p0 + i;
            const int l = p / kP;
            const float x = i ? xy.z : xy.x, y = i ? xy.w : xy.y, a = i ? aw.y : aw.x;
            int tok[4];
            float lh, lw;
            msda_sample_geom(x, y, lv.H[l], lv.W[l], lv.start[l], tok, lh, lw);
            int4 o;
            o.x = tok[0] >= 0 ? (tok[0] * M + m) * 32 : -1;
            o.y = tok[1] >= 0 ? (tok[1] * M + m) * 32 : -1;
            o.z = tok[2] >= 0 ? (tok[2] * M + m) * 32 : -1;
            o.w = tok[3] >= 0 ? (tok[3] * M + m) * 32 : -1;
            rec_off[p] = o;
            rec_f[p] = make_float4(lh, lw, a, 0.f);
          }
        }
      }
    }
    __syncwarp();

    const unsigned amask = __ballot_sync(0xffffffffu, active);
    if (active) {
      float go[C];
      {
        const float* gp = grad_out + qm * 32 + j * C;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
          const float4 t = ld_stream_f4(gp + c);
          go[c] = t.x; go[c + 1] = t.y; go[c + 2] = t.z; go[c + 3] = t.w;
        }
      }
      // results this lane will store: points p0 = 2j + 2G k (+0, +1)
      float keep_gx[Cfg::KP][2], keep_gy[Cfg::KP][2], keep_ga[Cfg::KP][2];
#pragma unroll
      for (int k = 0; k < Cfg::KP; ++k)
        for (int i = 0; i < 2; ++i) keep_gx[k][i] = keep_gy[k][i] = keep_ga[k][i] = 0.f;

#pragma unroll
      for (int p = 0; p < LP; ++p) {
        const int l = p / kP;
        const int4 o = rec_off[p];
        const float4 f = rec_f[p];
        const float lh = f.x, lw = f.y, a = f.z;
        const float hh = 1.f - lh, hw = 1.f - lw;
        RowFrag<C> r0, r1, r2, r3;
        row_load_or_zero(r0, value_b, o.x);
        row_load_or_zero(r1, value_b, o.y);
        row_load_or_zero(r2, value_b, o.z);
        row_load_or_zero(r3, value_b, o.w);
        // grad_value: corner weight x attention weight x grad_out   (cuh:125,134,143,152)
        const float t0 = a * (hh * hw), t1 = a * (hh * lw), t2 = a * (lh * hw), t3 = a * (lh * lw);
#pragma unroll
        for (int c = 0; kScatter && c < C; c += 4) {
          if (o.x >= 0) red_add_f4(gvalue_b + o.x + c, t0 * go[c], t0 * go[c + 1], t0 * go[c + 2], t0 * go[c + 3]);
          if (o.y >= 0) red_add_f4(gvalue_b + o.y + c, t1 * go[c], t1 * go[c + 1], t1 * go[c + 2], t1 * go[c + 3]);
          if (o.z >= 0) red_add_f4(gvalue_b + o.z + c, t2 * go[c], t2 * go[c + 1], t2 * go[c + 2], t2 * go[c + 3]);
          if (o.w >= 0) red_add_f4(gvalue_b + o.w + c, t3 * go[c], t3 * go[c + 1], t3 * go[c + 2], t3 * go[c + 3]);
        }
        // per-corner dot products with grad_out over this lane's channels
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          d0 = fmaf(go[c], r0.v[c], d0);
          d1 = fmaf(go[c], r1.v[c], d1);
          d2 = fmaf(go[c], r2.v[c], d2);
          d3 = fmaf(go[c], r3.v[c], d3);
        }
        // grad_attn_weight = sum_c grad_out * bilinear(value)                    (cuh:156)
        float ga = (hh * hw) * d0 + (hh * lw) * d1 + (lh * hw) * d2 + (lh * lw) * d3;
        // d/dw: -hh v00 + hh v01 - lh v10 + lh v11 ; d/dh: -hw v00 - lw v01 + hw v10 + lw v11
        float gx = (a * (float)lv.W[l]) * (hh * (d1 - d0) + lh * (d3 - d2));  // (cuh:157)
        float gy = (a * (float)lv.H[l]) * (hw * (d2 - d0) + lw * (d3 - d1));  // (cuh:158)
#pragma unroll
        for (int s = G / 2; s >= 1; s >>= 1) {
          ga += __shfl_xor_sync(amask, ga, s);
          gx += __shfl_xor_sync(amask, gx, s);
          gy += __shfl_xor_sync(amask, gy, s);
        }
        // route to the lane that decoded point p (compile-time indices after unrolling)
        if (j == (p % (2 * G)) / 2) {
          keep_gx[p / (2 * G)][p & 1] = gx;
          keep_gy[p / (2 * G)][p & 1] = gy;
          keep_ga[p / (2 * G)][p & 1] = ga;
        }
      }
#pragma unroll
      for (int k = 0; k < Cfg::KP; ++k) {
        const int p0 = 2 * j + 2 * G * k;
        if (p0 < LP) {
          st_stream_f4(grad_loc + qm * (LP * 2) + p0 * 2,
                       make_float4(keep_gx[k][0], keep_gy[k][0], keep_gx[k][1], keep_gy[k][1]));
          st_stream_f2(grad_attw + qm * LP + p0, make_float2(keep_ga[k][0], keep_ga[k][1]));
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace msda
