// msda_d32.cuh — tuned kernels for head_dim (channels) == 32; fp32 and bf16 value.
//
// Work decomposition (both directions)
//   grid  = (num_heads * tiles, batch); one thread block = ONE head x a tile of kTileQ
//           queries taken from `order` (patch-tiled for encoder self-attention, natural
//           order otherwise).  Heads sample in different directions, so a block that sticks
//           to one head keeps the value rows it gathers resident in L1.
//   warp  = 32/G "lane groups"; one lane group = one (query, head) pair.  Each lane covers
//           its share of the 32 channels of a value row with ONE 16-byte load:
//             fp32 value: G = 8 lanes x 4 channels  (row = 128 B = one full L1 line)
//             bf16 value: G = 4 lanes x 8 channels  (row =  64 B)
//   stage 1 lane j of a group decodes sampling points j, j+G, ...: bit-exact geometry
//           (msda_sample_geom) and a 16-byte record {row offset | corner mask, lh, lw, a}
//           per point, published to shared memory (conflict-free STS.128).
//   stage 2 all lanes of the group walk the L*P records (one broadcast LDS.128 per point),
//           rebuild the corner addresses / weights in registers, issue the four predicated
//           row gathers and blend (forward) or dot / scatter (backward).
// What binds them (ncu, profiles/): forward — the L1/shared data pipe (4 gathered rows per
// point are compulsory) together with instruction issue; backward — the chip-wide fp32
// reduction rate of L2 (REDG.E.ADD.F32x4, ~6.4 TB/s of payload measured in isolation).
//
// Algorithm restated from models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299
// (forward), :87-159 and :301-403 (backward); nothing is shared with that code's thread
// mapping (one thread per output channel, one-warp blocks, serial reductions).
#pragma once

#include <cuda_bf16.h>

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
#ifndef MSDA_TILE_Q
#define MSDA_TILE_Q 64
#endif
constexpr int kTileQ = MSDA_TILE_Q;  // queries per thread block

// ---- per-value-type traits ----------------------------------------------------------------
template <typename VT>
struct RowTraits;
template <>
struct RowTraits<float> {
  static constexpr int G = 8, C = 4;  // lanes per row, channels per lane
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[4]) {
    const float4 t = ld_stream_f4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store_stream(float* p, const float (&v)[4]) {
    st_stream_f4(p, make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <>
struct RowTraits<__nv_bfloat16> {
  static constexpr int G = 4, C = 8;
  static __device__ __forceinline__ void unpack(const uint4 t, float (&v)[8]) {
    // bf16 -> fp32 is a 16-bit shift: low half = even channel, high half = odd channel
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
    v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    unpack(__ldg(reinterpret_cast<const uint4*>(p)), v);
  }
  static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
    unpack(t, v);
  }
  static __device__ __forceinline__ void store_stream(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(v[0], v[1]); t.x = *reinterpret_cast<unsigned*>(&h);
    h = __floats2bfloat162_rn(v[2], v[3]); t.y = *reinterpret_cast<unsigned*>(&h);
    h = __floats2bfloat162_rn(v[4], v[5]); t.z = *reinterpret_cast<unsigned*>(&h);
    h = __floats2bfloat162_rn(v[6], v[7]); t.w = *reinterpret_cast<unsigned*>(&h);
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
  }
};

// ---- which channels a lane scatters into grad_value --------------------------------------------
// A row reduction is issued as 16-byte `red.global.add.v4.f32` instructions, and L2 performs it per
// 32-byte sector.  fp32 rows: the 8 lanes of a group hold 4 consecutive channels each, one instruction
// covers the 128-byte row in 4 full sectors.  bf16 rows: a lane holds 8 consecutive channels; scattering
// those as two instructions makes each instruction touch 4 HALF sectors (8 sector operations per row —
// measured: 7.9 M instead of 4.5 M reduction sectors per decoder layer).  So grad_out is re-dealt once per
// (query, head): instruction i of lane j covers channels 16 i + 4 j .. + 3, i.e. the 4 lanes of one
// instruction cover 64 contiguous bytes = 2 full sectors.
template <typename VT>
struct ScatterDeal;
template <>
struct ScatterDeal<float> {
  static constexpr int N = 1;
  static __device__ __forceinline__ void deal(const float (&go)[4], int, unsigned, float (&gs)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) gs[c] = go[c];
  }
  static __device__ __forceinline__ int offset(int j, int) { return 4 * j; }
};
template <>
struct ScatterDeal<__nv_bfloat16> {
  static constexpr int N = 2;
  static __device__ __forceinline__ void deal(const float (&go)[8], int j, unsigned mask, float (&gs)[8]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int src = 2 * i + (j >> 1);  // lane of the group that holds channels 16 i + 4 j .. + 3
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float lo = __shfl_sync(mask, go[c], src, 4);
        const float hi = __shfl_sync(mask, go[4 + c], src, 4);
        gs[4 * i + c] = (j & 1) ? hi : lo;
      }
    }
  }
  static __device__ __forceinline__ int offset(int j, int i) { return 16 * i + 4 * j; }
};

template <typename VT, int LP>
struct D32Cfg {
  static constexpr int G = RowTraits<VT>::G;
  static constexpr int C = RowTraits<VT>::C;
  static constexpr int GPW = 32 / G;                  // (query, head) pairs per warp
  static constexpr int QPP = kWarps * GPW;            // queries per pass of the block
  static constexpr int PASSES = kTileQ / QPP;
  static constexpr int KP = (LP + G - 1) / G;         // points decoded per lane
  static constexpr int REC_STRIDE = LP + 1;           // float4 units per lane group (+1: bank skew)
  static constexpr int SMEM_BYTES = kWarps * GPW * REC_STRIDE * 16;
  static_assert(kTileQ % QPP == 0, "tile must be a whole number of passes");
};

// Decodes this lane's points of (query, head) `qm` and publishes their records.
// record.x = element offset of corner (h0,w0)'s row within the image's value block
//            (a multiple of 32, possibly "virtual" when h0 or w0 is -1) | 4-bit corner mask
template <int G, int kL, int kP>
__device__ __forceinline__ void d32_decode_points(const float* __restrict__ loc,
                                                  const float* __restrict__ attw, const size_t qm,
                                                  const int j, const int m, const int M,
                                                  const MsdaLevels& lv, float4* rec,
                                                  const MsdaFused fz = MsdaFused{nullptr, 0}, const size_t bq = 0,
                                                  const unsigned gmask = 0xffffffffu, float* first_weight = nullptr) {
  constexpr int LP = kL * kP, KPTS = (LP + G - 1) / G;
  float2 xy[KPTS];
  float a[KPTS];
#pragma unroll
  for (int k = 0; k < KPTS; ++k) {
    const int p = j + G * k;
    xy[k] = make_float2(0.f, 0.f);
    a[k] = 0.f;
    if (p < LP) {
      xy[k] = ld_stream_f2(loc + (qm * LP + p) * 2);
      a[k] = ld_stream_f1(attw + qm * LP + p);
    }
  }
  if (fz.ref_dim) {
    // fused prologue: softmax over the (query, head)'s L*P logits, spread over the group's G lanes
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KPTS; ++k)
      if (j + G * k < LP) mx = fmaxf(mx, a[k]);
#pragma unroll
    for (int s = G / 2; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, s));
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KPTS; ++k) {
      a[k] = (j + G * k < LP) ? __expf(a[k] - mx) : 0.f;
      sum += a[k];
    }
#pragma unroll
    for (int s = G / 2; s >= 1; s >>= 1) sum += __shfl_xor_sync(gmask, sum, s);
#pragma unroll
    for (int k = 0; k < KPTS; ++k) {
      const int p = j + G * k;
      a[k] = a[k] / sum;
      if (p < LP) xy[k] = msda_fused_location(fz, bq, kL, p / kP, kP, lv.H[p / kP], lv.W[p / kP], xy[k]);
    }
  }
  if (first_weight) *first_weight = a[0];  // weight of point j (after the softmax, if fused)
#pragma unroll
  for (int k = 0; k < KPTS; ++k) {
    const int p = j + G * k;
    if (p < LP) {
      const int l = p / kP;
      const int W = lv.W[l];
      int tok[4];
      float lh, lw;
      const bool in = msda_sample_geom(xy[k].x, xy[k].y, lv.H[l], W, lv.start[l], tok, lh, lw, lv.coord_fma != 0);
      int base = 0, mask = 0;
      if (in) {
        // token of (h0, w0), valid or not: recover it from whichever corner exists
        const int t00 = tok[0] >= 0 ? tok[0] : tok[1] >= 0 ? tok[1] - 1 : tok[2] >= 0 ? tok[2] - W : tok[3] - W - 1;
        mask = (tok[0] >= 0) | ((tok[1] >= 0) << 1) | ((tok[2] >= 0) << 2) | ((tok[3] >= 0) << 3);
        base = mask ? (t00 * M + m) * 32 : 0;
      }
      rec[p] = make_float4(__int_as_float(base | mask), lh, lw, a[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// Resident blocks per SM the forward is compiled for.  Left to itself ptxas takes 56 registers (4 blocks, 46 % occupancy);
// measured per bs=2 encoder layer, fp32 / bf16: 1 -> 0.1374 / 0.1038 ms, 5 (48 regs) -> 0.1303 / 0.1035, 6 (40 regs) ->
// 0.1308 / 0.1020, 8 (32 regs, 4 bytes spilled) -> 0.1337 / 0.1112.
#ifndef MSDA_FWD_MINBLOCKS
#define MSDA_FWD_MINBLOCKS 6
#endif
template <typename VT, int kL, int kP, int kM>
__global__ void __launch_bounds__(kThreads, MSDA_FWD_MINBLOCKS)
msda_fwd_d32_kernel(const VT* __restrict__ value, const float* __restrict__ loc,
                    const float* __restrict__ attw, VT* __restrict__ out,
                    const int* __restrict__ order, const int order_len,
                    const __grid_constant__ MsdaLevels lv, const int S, const int M_rt, const int Lq,
                    const MsdaFused fz) {
  constexpr int LP = kL * kP;
  using Cfg = D32Cfg<VT, LP>;
  using RT = RowTraits<VT>;
  constexpr int G = Cfg::G, C = Cfg::C;
  extern __shared__ float4 smem[];
  const int M = kM ? kM : M_rt;  // kM > 0: number of heads known at compile time

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, j = lane % G;
  const int m = blockIdx.x % M;
  const int tile = blockIdx.x / M;
  const int b = blockIdx.y;
  const int M32 = M * 32;

  float4* rec = smem + (warp * Cfg::GPW + g) * Cfg::REC_STRIDE;
  const VT* value_b = value + (size_t)b * S * M32 + j * C;

#pragma unroll 1
  for (int pass = 0; pass < Cfg::PASSES; ++pass) {
    const int slot = tile * kTileQ + pass * Cfg::QPP + warp * Cfg::GPW + g;
    int q = -1;
    if (slot < order_len) q = order ? order[slot] : slot;
    const bool active = q >= 0;
    const size_t qm = ((size_t)b * Lq + (active ? q : 0)) * M + m;

    if (active)
      d32_decode_points<G, kL, kP>(loc, attw, qm, j, m, M, lv, rec, fz, (size_t)b * Lq + q,
                                   ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (g * G));
    __syncwarp();

    if (active) {
      float2 acc2[C / 2];  // channel pairs: the blends are packed FFMA2
#pragma unroll
      for (int c = 0; c < C / 2; ++c) acc2[c] = make_float2(0.f, 0.f);
#pragma unroll
      for (int p = 0; p < LP; ++p) {
        const int l = p / kP;
        const float4 r = rec[p];
        const int bm = __float_as_int(r.x);
        const float lh = r.y, lw = r.z, a = r.w;
        const float a_hh = a * (1.f - lh), a_lh = a * lh, hw = 1.f - lw;
        // one 64-bit address per row pair; the +M32 neighbour is an immediate when kM is known
        const VT* p0 = value_b + (ptrdiff_t)(bm & ~31);
        const VT* p2 = p0 + (ptrdiff_t)(lv.W[l] * M32);
        float v[C];
        if (bm & 1) {
          RT::load(p0, v);
          const float w = a_hh * hw;
          const float2 w2 = make_float2(w, w);
#pragma unroll
          for (int c = 0; c < C / 2; ++c) acc2[c] = ffma2(w2, make_float2(v[2 * c], v[2 * c + 1]), acc2[c]);
        }
        if (bm & 2) {
          RT::load(p0 + M32, v);
          const float w = a_hh * lw;
          const float2 w2 = make_float2(w, w);
#pragma unroll
          for (int c = 0; c < C / 2; ++c) acc2[c] = ffma2(w2, make_float2(v[2 * c], v[2 * c + 1]), acc2[c]);
        }
        if (bm & 4) {
          RT::load(p2, v);
          const float w = a_lh * hw;
          const float2 w2 = make_float2(w, w);
#pragma unroll
          for (int c = 0; c < C / 2; ++c) acc2[c] = ffma2(w2, make_float2(v[2 * c], v[2 * c + 1]), acc2[c]);
        }
        if (bm & 8) {
          RT::load(p2 + M32, v);
          const float w = a_lh * lw;
          const float2 w2 = make_float2(w, w);
#pragma unroll
          for (int c = 0; c < C / 2; ++c) acc2[c] = ffma2(w2, make_float2(v[2 * c], v[2 * c + 1]), acc2[c]);
        }
      }
      float acc[C];
#pragma unroll
      for (int c = 0; c < C / 2; ++c) { acc[2 * c] = acc2[c].x; acc[2 * c + 1] = acc2[c].y; }
      RT::store_stream(out + qm * 32 + j * C, acc);
    }
    __syncwarp();
  }
}

// Reduce-scatter of v[0..G) over the G lanes of a group: returns, in lane j, the sum over the
// group's lanes of v[j].  G-1 shuffles instead of the G*log2(G) of G separate butterflies.
template <int G>
__device__ __forceinline__ float group_reduce_scatter(float (&v)[G], const int j, const unsigned amask) {
#pragma unroll
  for (int h = G / 2; h >= 1; h >>= 1) {
    const bool hi = j & h;
#pragma unroll
    for (int k = 0; k < h; ++k) {
      const float send = hi ? v[k] : v[k + h];
      const float keep = hi ? v[k + h] : v[k];
      v[k] = keep + __shfl_xor_sync(amask, send, h);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------
// backward, atomic grad_value (REDG.E.ADD.F32x4); kScatter=false leaves grad_value alone
// (deterministic mode computes it separately, msda_det.cuh).  grad_out has value's type;
// all three gradients are fp32.
// ------------------------------------------------------------------------------------------
// (without the scatter the kernel is a pure gather: 64 registers / four blocks per SM measured faster,
// 0.181 ms against 0.200 ms per bs=2 encoder layer)
template <typename VT, int kL, int kP, int kM, bool kScatter>
__global__ void __launch_bounds__(kThreads, kScatter ? MSDA_BWD_MINBLOCKS : MSDA_BWD_MINBLOCKS + 1)
msda_bwd_d32_kernel(const VT* __restrict__ grad_out, const VT* __restrict__ value,
                    const float* __restrict__ loc, const float* __restrict__ attw,
                    float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_attw, const int* __restrict__ order,
                    const int order_len, const __grid_constant__ MsdaLevels lv, const int S,
                    const int M_rt, const int Lq) {
  constexpr int LP = kL * kP;
  using Cfg = D32Cfg<VT, LP>;
  using RT = RowTraits<VT>;
  constexpr int G = Cfg::G, C = Cfg::C;
  extern __shared__ float4 smem[];
  const int M = kM ? kM : M_rt;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, j = lane % G;
  const int m = blockIdx.x % M;
  const int tile = blockIdx.x / M;
  const int b = blockIdx.y;
  const int M32 = M * 32;

  float4* rec = smem + (warp * Cfg::GPW + g) * Cfg::REC_STRIDE;
  const size_t img = (size_t)b * S * M32 + j * C;
  const VT* value_b = value + img;
  float* gvalue_img = grad_value + (size_t)b * S * M32;
  using SD = ScatterDeal<VT>;

#pragma unroll 1
  for (int pass = 0; pass < Cfg::PASSES; ++pass) {
    const int slot = tile * kTileQ + pass * Cfg::QPP + warp * Cfg::GPW + g;
    int q = -1;
    if (slot < order_len) q = order ? order[slot] : slot;
    const bool active = q >= 0;
    const size_t qm = ((size_t)b * Lq + (active ? q : 0)) * M + m;

    if (active) d32_decode_points<G, kL, kP>(loc, attw, qm, j, m, M, lv, rec);
    __syncwarp();
    const unsigned amask = __ballot_sync(0xffffffffu, active);

    if (active) {
      float go[C], gs[C];
      RT::load_stream(grad_out + qm * 32 + j * C, go);
      if (kScatter) SD::deal(go, j, amask, gs);
      // points are reduced in blocks of G: lane j ends up owning point (G*blk + j) — the same
      // point it decoded, so it also stores that point's gradients.
#pragma unroll
      for (int blk = 0; blk < Cfg::KP; ++blk) {
        float pgx[G], pgy[G], pga[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int p = blk * G + i;
          pgx[i] = pgy[i] = pga[i] = 0.f;
          if (p < LP) {
            const int l = p / kP;
            const float4 r = rec[p];
            const int bm = __float_as_int(r.x);
            const float lh = r.y, lw = r.z, a = r.w;
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float a_hh = a * hh, a_lh = a * lh;
            const ptrdiff_t o0 = (ptrdiff_t)(bm & ~31);
            const ptrdiff_t o2 = o0 + (ptrdiff_t)(lv.W[l] * M32);
            // per corner: gather the row, scatter corner weight x attention weight x grad_out into
            // grad_value (cuh:125,134,143,152), and dot the row with grad_out over this lane's channels
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (bm & (1 << k)) {
                const ptrdiff_t o = ((k & 2) ? o2 : o0) + ((k & 1) ? M32 : 0);
                float v[C];
                RT::load(value_b + o, v);
                if (kScatter) {
                  const float t = ((k & 2) ? a_lh : a_hh) * ((k & 1) ? lw : hw);
#pragma unroll
                  for (int i2 = 0; i2 < SD::N; ++i2)
                    red_add_f4(gvalue_img + o + SD::offset(j, i2), t * gs[4 * i2], t * gs[4 * i2 + 1],
                               t * gs[4 * i2 + 2], t * gs[4 * i2 + 3]);
                }
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) s = fmaf(go[c], v[c], s);
                d[k] = s;
              }
            }
            // grad_attn_weight = sum_c grad_out * bilinear(value)                      (cuh:156)
            pga[i] = hh * (hw * d[0] + lw * d[1]) + lh * (hw * d[2] + lw * d[3]);
            // d/dx: W*a*(-hh v00 + hh v01 - lh v10 + lh v11); d/dy: H*a*(-hw v00 - lw v01 + hw v10 + lw v11)
            pgx[i] = (a * (float)lv.W[l]) * (hh * (d[1] - d[0]) + lh * (d[3] - d[2]));  // (cuh:157)
            pgy[i] = (a * (float)lv.H[l]) * (hw * (d[2] - d[0]) + lw * (d[3] - d[1]));  // (cuh:158)
          }
        }
        const float gx = group_reduce_scatter<G>(pgx, j, amask);
        const float gy = group_reduce_scatter<G>(pgy, j, amask);
        const float ga = group_reduce_scatter<G>(pga, j, amask);
        const int p = blk * G + j;
        if (p < LP) {
          st_stream_f2(grad_loc + (qm * LP + p) * 2, make_float2(gx, gy));
          st_stream_f1(grad_attw + qm * LP + p, ga);
        }
      }
    }
    __syncwarp();
  }
}

// ==========================================================================================
// "split" variants for small problems (decoder cross-attention: a few thousand queries).
// With one lane group per (query, head) the grid is too small to hide the latency of 16
// dependent gather rounds, so here a whole WARP owns one (query, head): lane group g handles
// points g, g+GPW, ... (4x / 8x more rows in flight per query), and the forward combines the
// groups' partial sums with xor-shuffles.  Block = 4 warps = 4 consecutive queries of one head.
// ==========================================================================================
constexpr int kSplitThreads = 128;

template <typename VT, int kL, int kP, int kM>
__global__ void __launch_bounds__(kSplitThreads)
msda_fwd_d32_split_kernel(const VT* __restrict__ value, const float* __restrict__ loc,
                          const float* __restrict__ attw, VT* __restrict__ out,
                          const __grid_constant__ MsdaLevels lv, const int S, const int M_rt, const int Lq,
                          const MsdaFused fz) {
  constexpr int LP = kL * kP;
  using RT = RowTraits<VT>;
  constexpr int G = RT::G, C = RT::C, GPW = 32 / G;
  constexpr int NPG = (LP + GPW - 1) / GPW;  // points per lane group
  __shared__ float4 smem[(kSplitThreads / 32) * (LP + 1)];
  const int M = kM ? kM : M_rt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, j = lane % G;
  const int m = blockIdx.x % M;
  const int q = (blockIdx.x / M) * (kSplitThreads / 32) + warp;
  const int b = blockIdx.y;
  const int M32 = M * 32;
  if (q >= Lq) return;  // warp-uniform
  float4* rec = smem + warp * (LP + 1);
  const VT* value_b = value + (size_t)b * S * M32 + j * C;
  const size_t qm = ((size_t)b * Lq + q) * M + m;
  d32_decode_points<32, kL, kP>(loc, attw, qm, lane, m, M, lv, rec, fz, (size_t)b * Lq + q);
  __syncwarp();
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
  for (int i = 0; i < NPG; ++i) {
    const int p = g + GPW * i;
    if (p < LP) {
      const int l = p / kP;
      const float4 r = rec[p];
      const int bm = __float_as_int(r.x);
      const float lh = r.y, lw = r.z, a = r.w;
      const float a_hh = a * (1.f - lh), a_lh = a * lh, hw = 1.f - lw;
      const VT* p0 = value_b + (ptrdiff_t)(bm & ~31);
      const VT* p2 = p0 + (ptrdiff_t)(lv.W[l] * M32);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (bm & (1 << k)) {
          float v[C];
          RT::load(((k & 2) ? p2 : p0) + ((k & 1) ? M32 : 0), v);
          const float w = ((k & 2) ? a_lh : a_hh) * ((k & 1) ? lw : hw);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(w, v[c], acc[c]);
        }
      }
    }
  }
#pragma unroll
  for (int s = G; s < 32; s <<= 1)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], s);
  if (g == 0) RT::store_stream(out + qm * 32 + j * C, acc);
}

// Resident blocks per SM (of 128 threads) the split backward is compiled for.  Measured on the graph-replayed decoder step
// (6 layers): fp32 value 1 -> 0.3185 ms, 12 (40 registers) -> 0.3115 ms, 16 (32 registers, spills) -> 0.3194 ms; bf16 value
// 1 (54 registers) -> 0.2931 ms, 12 -> 0.3079 ms, 16 -> 0.3427 ms.  Hence 12 for fp32 and ptxas's own choice (no minimum) for bf16.
#ifndef MSDA_SPLIT_BWD_MINBLOCKS
#define MSDA_SPLIT_BWD_MINBLOCKS(VT) (sizeof(VT) == 4 ? 12 : 0)  // 0 = no minimum
#endif
template <typename VT, int kL, int kP, int kM, bool kScatter>
__global__ void __launch_bounds__(kSplitThreads, MSDA_SPLIT_BWD_MINBLOCKS(VT))
msda_bwd_d32_split_kernel(const VT* __restrict__ grad_out, const VT* __restrict__ value,
                          const float* __restrict__ loc, const float* __restrict__ attw,
                          float* __restrict__ grad_value, float* __restrict__ grad_loc,
                          float* __restrict__ grad_attw, const __grid_constant__ MsdaLevels lv,
                          const int S, const int M_rt, const int Lq, const MsdaFused fz) {
  constexpr int LP = kL * kP;
  using RT = RowTraits<VT>;
  constexpr int G = RT::G, C = RT::C, GPW = 32 / G;
  constexpr int NPG = (LP + GPW - 1) / GPW;
  __shared__ float4 smem[(kSplitThreads / 32) * (LP + 1)];
  __shared__ float4 res_s[(kSplitThreads / 32) * LP];  // fused prologue: (d/dx, d/dy, d/da) of every point of the warp's (query, head)
  const int M = kM ? kM : M_rt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, j = lane % G;
  const int m = blockIdx.x % M;
  const int q = (blockIdx.x / M) * (kSplitThreads / 32) + warp;
  const int b = blockIdx.y;
  const int M32 = M * 32;
  if (q >= Lq) return;
  float4* rec = smem + warp * (LP + 1);
  const size_t img = (size_t)b * S * M32 + j * C;
  const VT* value_b = value + img;
  float* gvalue_img = grad_value + (size_t)b * S * M32;
  using SD = ScatterDeal<VT>;
  const size_t qm = ((size_t)b * Lq + q) * M + m;
  const bool fused = fz.ref_dim != 0;  // grid-uniform
  float4* res = res_s + warp * LP;
  float my_a = 0.f;  // lane p < LP: attention weight of point p
  d32_decode_points<32, kL, kP>(loc, attw, qm, lane, m, M, lv, rec, fz, (size_t)b * Lq + q, 0xffffffffu, &my_a);
  __syncwarp();
  float go[C], gs[C];
  RT::load_stream(grad_out + qm * 32 + j * C, go);
  if (kScatter) SD::deal(go, j, 0xffffffffu, gs);
#pragma unroll
  for (int i = 0; i < NPG; ++i) {
    const int p = g + GPW * i;
    float ga = 0.f, gx = 0.f, gy = 0.f;
    if (p < LP) {
      const int l = p / kP;
      const float4 r = rec[p];
      const int bm = __float_as_int(r.x);
      const float lh = r.y, lw = r.z, a = r.w;
      const float hh = 1.f - lh, hw = 1.f - lw;
      const ptrdiff_t o0 = (ptrdiff_t)(bm & ~31);
      const ptrdiff_t o2 = o0 + (ptrdiff_t)(lv.W[l] * M32);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (bm & (1 << k)) {
          const ptrdiff_t o = ((k & 2) ? o2 : o0) + ((k & 1) ? M32 : 0);
          float v[C];
          RT::load(value_b + o, v);
          float sdot = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) sdot = fmaf(go[c], v[c], sdot);
          d[k] = sdot;
        }
      }
      ga = hh * (hw * d[0] + lw * d[1]) + lh * (hw * d[2] + lw * d[3]);
      gx = (a * (float)lv.W[l]) * (hh * (d[1] - d[0]) + lh * (d[3] - d[2]));
      gy = (a * (float)lv.H[l]) * (hw * (d[2] - d[0]) + lw * (d[3] - d[1]));
    }
#pragma unroll
    for (int s = G / 2; s >= 1; s >>= 1) {
      ga += __shfl_xor_sync(0xffffffffu, ga, s);
      gx += __shfl_xor_sync(0xffffffffu, gx, s);
      gy += __shfl_xor_sync(0xffffffffu, gy, s);
    }
    if (j == 0 && p < LP) {
      if (fused) {
        res[p] = make_float4(gx, gy, ga, 0.f);
      } else {
        st_stream_f2(grad_loc + (qm * LP + p) * 2, make_float2(gx, gy));
        st_stream_f1(grad_attw + qm * LP + p, ga);
      }
    }
  }
  if (fused) {
    // Fused prologue (ms_deform_attn.py:98-111 backwards): lane p turns point p's gradients with respect to the
    // sampling location and the attention weight into those of the RAW offset and logit.  Softmax backward:
    // dL/dlogit_p = a_p * (dL/da_p - sum_k a_k * dL/da_k), the sum over the (query, head)'s L*P points (a skipped
    // sample has dL/da = 0 but keeps its softmax weight); location: loc = ref + off / (W, H)  |  ref.xy + off / P *
    // ref.wh * 0.5, divisions and products in autograd's order.  Coalesced stores: 128 + 64 contiguous bytes per warp.
    __syncwarp();
    const int p = lane, l = min(p / kP, kL - 1);
    const float4 r = p < LP ? res[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    float dot = p < LP ? my_a * r.z : 0.f;
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, sft);
    if (p < LP) {
      const float* rp = fz.ref + (((size_t)b * Lq + q) * kL + l) * fz.ref_dim;
      const float gx = fz.ref_dim == 2 ? r.x / (float)lv.W[l] : r.x * 0.5f * rp[2] / (float)kP;
      const float gy = fz.ref_dim == 2 ? r.y / (float)lv.H[l] : r.y * 0.5f * rp[3] / (float)kP;
      st_stream_f2(grad_loc + (qm * LP + p) * 2, make_float2(gx, gy));
      st_stream_f1(grad_attw + qm * LP + p, my_a * (r.z - dot));
    }
  }
  if (kScatter) {
    // the scatter needs no value rows, only the records and grad_out: it runs after the gather phase, behind the
    // completion of the preceding kernel on the stream (the zero-fill of grad_value when the launch allowed an early
    // start; otherwise the wait returns at once)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NPG; ++i) {
      const int p = g + GPW * i;
      if (p < LP) {
        const int l = p / kP;
        const float4 r = rec[p];
        const int bm = __float_as_int(r.x);
        const float lh = r.y, lw = r.z, a = r.w;
        const float a_hh = a * (1.f - lh), a_lh = a * lh, hw = 1.f - lw;
        const ptrdiff_t o0 = (ptrdiff_t)(bm & ~31);
        const ptrdiff_t o2 = o0 + (ptrdiff_t)(lv.W[l] * M32);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (bm & (1 << k)) {
            const ptrdiff_t o = ((k & 2) ? o2 : o0) + ((k & 1) ? M32 : 0);
            const float t = ((k & 2) ? a_lh : a_hh) * ((k & 1) ? lw : hw);
#pragma unroll
            for (int i2 = 0; i2 < SD::N; ++i2)
              red_add_f4(gvalue_img + o + SD::offset(j, i2), t * gs[4 * i2], t * gs[4 * i2 + 1], t * gs[4 * i2 + 2],
                         t * gs[4 * i2 + 3]);
          }
        }
      }
    }
  }
}

}  // namespace msda
