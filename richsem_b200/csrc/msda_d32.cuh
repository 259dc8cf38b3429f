// msda_d32.cuh — tuned fp32 kernels for head_dim (channels) == 32.
//
// Work decomposition (both directions)
//   grid  = (num_heads * tiles, batch); one thread block = ONE head x a tile of kTileQ
//           queries taken from `order` (patch-tiled for encoder self-attention, natural
//           order otherwise).  Heads sample in different directions, so a block that sticks
//           to one head keeps the value rows it gathers resident in L1.
//   warp  = four "lane groups" of 8 lanes; one lane group = one (query, head) pair.  The 8
//           lanes cover the 32 channels of a value row with one LDG.128 each, so a gathered
//           row is always one full 128-byte line = one L1 wavefront.
//   stage 1 lane j of a group decodes sampling points j, j+8, ...: bit-exact geometry
//           (msda_sample_geom) and a 16-byte record {row offset | corner mask, lh, lw, a}
//           per point, published to shared memory (conflict-free STS.128).
//   stage 2 all lanes of the group walk the L*P records (one broadcast LDS.128 per point),
//           rebuild the four corner offsets / weights in registers, issue the four
//           predicated row gathers and blend.
// The kernels are bound by the L1/shared data pipe (4 gather wavefronts per point are
// compulsory for fp32 rows), so everything else is organised to spend as few extra
// wavefronts as possible: 16-byte records, shuffle-light reductions in the backward.
//
// Algorithm restated from models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299
// (forward), :87-159 and :301-403 (backward); nothing is shared with that code's thread
// mapping (one thread per output channel, one-warp blocks, serial reductions).
#pragma once

#include "msda_common.cuh"

namespace msda {

#ifndef MSDA_THREADS
#define MSDA_THREADS 256
#endif
#ifndef MSDA_BWD_MINBLOCKS
#define MSDA_BWD_MINBLOCKS 3
#endif
constexpr int kThreads = MSDA_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kTileQ = 64;  // queries per thread block
constexpr int kG = 8;       // lanes per (query, head)
constexpr int kC = 4;       // channels per lane

template <int LP>
struct D32Cfg {
  static constexpr int GPW = 32 / kG;                 // (query, head) pairs per warp
  static constexpr int QPP = kWarps * GPW;            // queries per pass of the block
  static constexpr int PASSES = kTileQ / QPP;
  static constexpr int KP = (LP + kG - 1) / kG;       // points decoded per lane
  static constexpr int REC_STRIDE = LP + 1;           // float4 units per lane group (+1: bank skew)
  static constexpr int SMEM_BYTES = kWarps * GPW * REC_STRIDE * 16;
  static_assert(kTileQ % QPP == 0, "tile must be a whole number of passes");
};

// Decodes this lane's points of (query, head) `qm` and publishes their records.
// record.x = element offset of corner (h0,w0)'s row within the image's value block
//            (a multiple of 32, possibly "virtual" when h0 or w0 is -1) | 4-bit corner mask
template <int kL, int kP>
__device__ __forceinline__ void d32_decode_points(const float* __restrict__ loc,
                                                  const float* __restrict__ attw, const size_t qm,
                                                  const int j, const int m, const int M,
                                                  const MsdaLevels& lv, float4* rec) {
  constexpr int LP = kL * kP;
#pragma unroll
  for (int k = 0; k < D32Cfg<LP>::KP; ++k) {
    const int p = j + kG * k;
    if (p < LP) {
      const int l = p / kP;
      const float2 xy = ld_stream_f2(loc + (qm * LP + p) * 2);
      const float a = ld_stream_f1(attw + qm * LP + p);
      const int W = lv.W[l];
      int tok[4];
      float lh, lw;
      const bool in = msda_sample_geom(xy.x, xy.y, lv.H[l], W, lv.start[l], tok, lh, lw);
      int base = 0, mask = 0;
      if (in) {
        // token of (h0, w0), valid or not: recover it from whichever corner exists
        const int t00 = tok[0] >= 0 ? tok[0] : tok[1] >= 0 ? tok[1] - 1 : tok[2] >= 0 ? tok[2] - W : tok[3] - W - 1;
        mask = (tok[0] >= 0) | ((tok[1] >= 0) << 1) | ((tok[2] >= 0) << 2) | ((tok[3] >= 0) << 3);
        base = mask ? (t00 * M + m) * 32 : 0;
      }
      rec[p] = make_float4(__int_as_float(base | mask), lh, lw, a);
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int kL, int kP, int kM>
__global__ void __launch_bounds__(kThreads)
msda_fwd_d32_kernel(const float* __restrict__ value, const float* __restrict__ loc,
                    const float* __restrict__ attw, float* __restrict__ out,
                    const int* __restrict__ order, const int order_len,
                    const __grid_constant__ MsdaLevels lv, const int S, const int M_rt, const int Lq) {
  constexpr int LP = kL * kP;
  using Cfg = D32Cfg<LP>;
  extern __shared__ float4 smem[];
  const int M = kM ? kM : M_rt;  // kM > 0: number of heads known at compile time

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / kG, j = lane % kG;
  const int m = blockIdx.x % M;
  const int tile = blockIdx.x / M;
  const int b = blockIdx.y;
  const int M32 = M * 32;

  float4* rec = smem + (warp * Cfg::GPW + g) * Cfg::REC_STRIDE;
  const float* value_b = value + (size_t)b * S * M32 + j * kC;

#pragma unroll 1
  for (int pass = 0; pass < Cfg::PASSES; ++pass) {
    const int slot = tile * kTileQ + pass * Cfg::QPP + warp * Cfg::GPW + g;
    int q = -1;
    if (slot < order_len) q = order ? order[slot] : slot;
    const bool active = q >= 0;
    const size_t qm = ((size_t)b * Lq + (active ? q : 0)) * M + m;

    if (active) d32_decode_points<kL, kP>(loc, attw, qm, j, m, M, lv, rec);
    __syncwarp();

    if (active) {
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
      for (int p = 0; p < LP; ++p) {
        const int l = p / kP;
        const float4 r = rec[p];
        const int bm = __float_as_int(r.x);
        const float lh = r.y, lw = r.z, a = r.w;
        const float a_hh = a * (1.f - lh), a_lh = a * lh, hw = 1.f - lw;
        // one 64-bit address per row pair; the +M32 neighbour is an immediate when kM is known
        const float* p0 = value_b + (ptrdiff_t)(bm & ~31);
        const float* p2 = p0 + (ptrdiff_t)(lv.W[l] * M32);
        if (bm & 1) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p0));
          const float w = a_hh * hw;
          acc0 = fmaf(w, v.x, acc0); acc1 = fmaf(w, v.y, acc1); acc2 = fmaf(w, v.z, acc2); acc3 = fmaf(w, v.w, acc3);
        }
        if (bm & 2) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p0 + M32));
          const float w = a_hh * lw;
          acc0 = fmaf(w, v.x, acc0); acc1 = fmaf(w, v.y, acc1); acc2 = fmaf(w, v.z, acc2); acc3 = fmaf(w, v.w, acc3);
        }
        if (bm & 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p2));
          const float w = a_lh * hw;
          acc0 = fmaf(w, v.x, acc0); acc1 = fmaf(w, v.y, acc1); acc2 = fmaf(w, v.z, acc2); acc3 = fmaf(w, v.w, acc3);
        }
        if (bm & 8) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p2 + M32));
          const float w = a_lh * lw;
          acc0 = fmaf(w, v.x, acc0); acc1 = fmaf(w, v.y, acc1); acc2 = fmaf(w, v.z, acc2); acc3 = fmaf(w, v.w, acc3);
        }
      }
      st_stream_f4(out + qm * 32 + j * kC, make_float4(acc0, acc1, acc2, acc3));
    }
    __syncwarp();
  }
}

// Reduce-scatter of v[0..7] over the 8 lanes of a group: returns, in lane j, the sum over the
// group's lanes of v[j].  7 shuffles instead of the 24 that 8 separate butterflies would take.
__device__ __forceinline__ float group_reduce_scatter8(const float (&v)[8], const int j,
                                                       const unsigned amask) {
  float a[4], b2[2];
  const bool hi4 = j & 4, hi2 = j & 2, hi1 = j & 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = hi4 ? v[k] : v[k + 4];
    const float keep = hi4 ? v[k + 4] : v[k];
    a[k] = keep + __shfl_xor_sync(amask, send, 4);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = hi2 ? a[k] : a[k + 2];
    const float keep = hi2 ? a[k + 2] : a[k];
    b2[k] = keep + __shfl_xor_sync(amask, send, 2);
  }
  const float send = hi1 ? b2[0] : b2[1];
  const float keep = hi1 ? b2[1] : b2[0];
  return keep + __shfl_xor_sync(amask, send, 1);
}

// ------------------------------------------------------------------------------------------
// backward, atomic grad_value (REDG.E.ADD.F32x4); kScatter=false leaves grad_value alone
// (deterministic mode computes it separately, msda_det.cuh)
// ------------------------------------------------------------------------------------------
template <int kL, int kP, int kM, bool kScatter>
__global__ void __launch_bounds__(kThreads, MSDA_BWD_MINBLOCKS)
msda_bwd_d32_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                    const float* __restrict__ loc, const float* __restrict__ attw,
                    float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_attw, const int* __restrict__ order,
                    const int order_len, const __grid_constant__ MsdaLevels lv, const int S,
                    const int M_rt, const int Lq) {
  constexpr int LP = kL * kP;
  using Cfg = D32Cfg<LP>;
  extern __shared__ float4 smem[];
  const int M = kM ? kM : M_rt;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / kG, j = lane % kG;
  const int m = blockIdx.x % M;
  const int tile = blockIdx.x / M;
  const int b = blockIdx.y;
  const int M32 = M * 32;

  float4* rec = smem + (warp * Cfg::GPW + g) * Cfg::REC_STRIDE;
  const float* value_b = value + (size_t)b * S * M32 + j * kC;
  const ptrdiff_t gdelta = grad_value - value;  // same element offsets in value and grad_value

#pragma unroll 1
  for (int pass = 0; pass < Cfg::PASSES; ++pass) {
    const int slot = tile * kTileQ + pass * Cfg::QPP + warp * Cfg::GPW + g;
    int q = -1;
    if (slot < order_len) q = order ? order[slot] : slot;
    const bool active = q >= 0;
    const size_t qm = ((size_t)b * Lq + (active ? q : 0)) * M + m;

    if (active) d32_decode_points<kL, kP>(loc, attw, qm, j, m, M, lv, rec);
    __syncwarp();
    const unsigned amask = __ballot_sync(0xffffffffu, active);

    if (active) {
      const float4 go = ld_stream_f4(grad_out + qm * 32 + j * kC);
      // points are reduced in blocks of 8: lane j ends up owning point (8*blk + j) — the same
      // point it decoded, so it also stores that point's gradients.
#pragma unroll
      for (int blk = 0; blk < Cfg::KP; ++blk) {
        float pgx[8], pgy[8], pga[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = blk * 8 + i;
          pgx[i] = pgy[i] = pga[i] = 0.f;
          if (p < LP) {
            const int l = p / kP;
            const float4 r = rec[p];
            const int bm = __float_as_int(r.x);
            const float lh = r.y, lw = r.z, a = r.w;
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float a_hh = a * hh, a_lh = a * lh;
            const float* p0 = value_b + (ptrdiff_t)(bm & ~31);
            const float* p2 = p0 + (ptrdiff_t)(lv.W[l] * M32);
            // per corner: gather the row, scatter corner weight x attention weight x grad_out into
            // grad_value (cuh:125,134,143,152), and dot the row with grad_out over this lane's channels
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
            if (bm & 1) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(p0));
              if (kScatter) { const float t = a_hh * hw; red_add_f4(const_cast<float*>(p0) + gdelta, t * go.x, t * go.y, t * go.z, t * go.w); }
              d0 = go.x * v.x + go.y * v.y + go.z * v.z + go.w * v.w;
            }
            if (bm & 2) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(p0 + M32));
              if (kScatter) { const float t = a_hh * lw; red_add_f4(const_cast<float*>(p0) + gdelta + M32, t * go.x, t * go.y, t * go.z, t * go.w); }
              d1 = go.x * v.x + go.y * v.y + go.z * v.z + go.w * v.w;
            }
            if (bm & 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(p2));
              if (kScatter) { const float t = a_lh * hw; red_add_f4(const_cast<float*>(p2) + gdelta, t * go.x, t * go.y, t * go.z, t * go.w); }
              d2 = go.x * v.x + go.y * v.y + go.z * v.z + go.w * v.w;
            }
            if (bm & 8) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(p2 + M32));
              if (kScatter) { const float t = a_lh * lw; red_add_f4(const_cast<float*>(p2) + gdelta + M32, t * go.x, t * go.y, t * go.z, t * go.w); }
              d3 = go.x * v.x + go.y * v.y + go.z * v.z + go.w * v.w;
            }
            // grad_attn_weight = sum_c grad_out * bilinear(value)                      (cuh:156)
            pga[i] = hh * (hw * d0 + lw * d1) + lh * (hw * d2 + lw * d3);
            // d/dx: W*a*(-hh v00 + hh v01 - lh v10 + lh v11); d/dy: H*a*(-hw v00 - lw v01 + hw v10 + lw v11)
            pgx[i] = (a * (float)lv.W[l]) * (hh * (d1 - d0) + lh * (d3 - d2));  // (cuh:157)
            pgy[i] = (a * (float)lv.H[l]) * (hw * (d2 - d0) + lw * (d3 - d1));  // (cuh:158)
          }
        }
        const float gx = group_reduce_scatter8(pgx, j, amask);
        const float gy = group_reduce_scatter8(pgy, j, amask);
        const float ga = group_reduce_scatter8(pga, j, amask);
        const int p = blk * 8 + j;
        if (p < LP) {
          st_stream_f2(grad_loc + (qm * LP + p) * 2, make_float2(gx, gy));
          st_stream_f1(grad_attw + qm * LP + p, ga);
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace msda
