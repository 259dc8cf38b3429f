// msda_d32_agg.cuh — backward for head_dim 32 with on-chip pre-aggregation of grad_value.
//
// Why: the plain backward (msda_d32.cuh) sends one 128-byte fp32 reduction per sampled corner
// to L2 — 22.75 M of them per bs=2 encoder layer — and L2 retires scattered fp32 row
// reductions at ~6.4 TB/s chip-wide (scratch/tma_red.cu), which alone is 455 us.  Encoder
// self-attention is spatially local: the 64 queries of a patch-ordered block put their
// 64*L*P*4 = 4096 corner contributions on only a few hundred distinct value rows.  Shared-memory
// float atomics are a CAS loop on sm_100a (14.7 cycles per row, scratch/smem_atomics.cu), so the
// merge is done "owner computes":
//
//   phase 1  the block decodes all 64 x L*P sampling points up front (4 points of one level per
//            thread), finds each level's window origin (min h0, min w0 over the block), and gives
//            every corner that falls inside the level's 20x20-cell window a SLOT (level, dh, dw).
//            A shared-memory histogram over the slots (native int32 ATOMS.ADD) hands each
//            contribution its rank inside the slot; an in-place exclusive scan turns counts into
//            offsets; the decode threads then write their {slot, query, coefficient} entries
//            straight to their sorted positions — a counting sort with one pass over the data.
//   phase 2  the usual gather (msda_d32.cuh stage 2) for grad_sampling_loc / grad_attn_weight;
//            it also parks each query's grad_out row in shared memory.  Points outside the
//            window ("direct" bit in the record) fall back to global reductions.
//   phase 3  the sorted list is cut into equal chunks, one per lane group; a group walks its
//            chunk accumulating coef * grad_out[q] in registers and issues ONE global reduction
//            per run of equal slots.  A slot cut by a chunk boundary is simply reduced twice.
//
// Gradient formulas: models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159.
#pragma once

#include <climits>

#include "msda_d32.cuh"

namespace msda {

constexpr int kAggW = 20;                  // window side, in cells, per level
constexpr int kAggSlots = kAggW * kAggW;   // slots per level
constexpr int kAggThreads = 256;

template <int kL>
struct AggCfg {
  static constexpr int LP = kL * 4;
  static constexpr int NLV = (kL + 3) / 4;                   // levels decoded per thread
  static constexpr int NSLOT = kL * kAggSlots;
  static constexpr int SPT = (NSLOT + kAggThreads - 1) / kAggThreads;  // slots per thread in the scan
  static constexpr int NSLOT_PAD = SPT * kAggThreads;
  static constexpr int REC_STRIDE = LP + 1;                  // float4 per query (+1: bank skew)
  static constexpr int REC_BYTES = kTileQ * REC_STRIDE * 16;
  static constexpr int CONTRIB_BYTES = kTileQ * LP * 4 * 8;
  static constexpr int GO_BYTES = kTileQ * 32 * 4;
  static constexpr int HIST_BYTES = NSLOT_PAD * 4;
  static constexpr int MISC_INTS = 32 + kTileQ;              // origins, warp totals, total, query ids
  static constexpr int SMEM_BYTES = REC_BYTES + CONTRIB_BYTES + GO_BYTES + HIST_BYTES + MISC_INTS * 4;
  static_assert(kL <= 8, "window origins are kept in 8-entry arrays");
  static_assert(NSLOT < (1 << 18) && kTileQ * LP * 4 <= 8192, "slot id / rank must fit the packed fields");
};

template <typename VT, int kL, int kM>
__global__ void __launch_bounds__(kAggThreads, 2)
msda_bwd_d32_agg_kernel(const VT* __restrict__ grad_out, const VT* __restrict__ value,
                        const float* __restrict__ loc, const float* __restrict__ attw,
                        float* __restrict__ grad_value, float* __restrict__ grad_loc,
                        float* __restrict__ grad_attw, const int* __restrict__ order,
                        const int order_len, const __grid_constant__ MsdaLevels lv, const int S,
                        const int M_rt, const int Lq) {
  using Cfg = AggCfg<kL>;
  using RT = RowTraits<VT>;
  constexpr int kP = 4, LP = Cfg::LP, NLV = Cfg::NLV;
  constexpr int G = RT::G, C = RT::C, GPW = 32 / G;
  constexpr int QPP = (kAggThreads / 32) * GPW, PASSES = kTileQ / QPP, KPG = (LP + G - 1) / G;
  static_assert(kTileQ % QPP == 0 && kTileQ * 4 == kAggThreads, "decode maps 4 threads to a query");

  extern __shared__ __align__(16) unsigned char smraw[];
  float4* rec = reinterpret_cast<float4*>(smraw);
  int2* contrib = reinterpret_cast<int2*>(smraw + Cfg::REC_BYTES);
  float* go_s = reinterpret_cast<float*>(smraw + Cfg::REC_BYTES + Cfg::CONTRIB_BYTES);
  int* hist = reinterpret_cast<int*>(smraw + Cfg::REC_BYTES + Cfg::CONTRIB_BYTES + Cfg::GO_BYTES);
  int* misc = hist + Cfg::NSLOT_PAD;  // [0,8) hmin  [8,16) wmin  [16,24) warp totals  [24] total
  int* qidx = misc + 32;              // query index of each of the block's 64 slots (-1: none)

  const int M = kM ? kM : M_rt;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int m = blockIdx.x % M;
  const int tile = blockIdx.x / M;
  const int b = blockIdx.y;
  const int M32 = M * 32;
  const bool fma = lv.coord_fma != 0;

  // ---- phase 0 ---------------------------------------------------------------------------
  for (int i = t; i < Cfg::NSLOT_PAD; i += kAggThreads) hist[i] = 0;
  if (t < 16) misc[t] = INT_MAX;
  __syncthreads();

  // ---- phase 1a: decode; thread = (query ql, level residue r) --------------------------------
  const int ql = t >> 2, r = t & 3;
  int q = -1;
  {
    const int oslot = tile * kTileQ + ql;
    if (oslot < order_len) q = order ? order[oslot] : oslot;
  }
  const bool act = q >= 0;
  const size_t qm = ((size_t)b * Lq + (act ? q : 0)) * M + m;
  if (r == 0) qidx[ql] = q;

  int p_bm[NLV][4], p_h0[NLV][4], p_w0[NLV][4];
  float p_lh[NLV][4], p_lw[NLV][4], p_a[NLV][4];
#pragma unroll
  for (int li = 0; li < NLV; ++li) {
    const int l = r + 4 * li;
    int hmn = INT_MAX, wmn = INT_MAX;
#pragma unroll
    for (int i = 0; i < 4; ++i) { p_bm[li][i] = 0; p_h0[li][i] = 0; p_w0[li][i] = 0; p_lh[li][i] = p_lw[li][i] = p_a[li][i] = 0.f; }
    if (l < kL && act) {
      const float* lp = loc + (qm * LP + l * kP) * 2;
      const float4 xy01 = ld_stream_f4(lp), xy23 = ld_stream_f4(lp + 4);
      const float4 aw = ld_stream_f4(attw + qm * LP + l * kP);
      const int H = lv.H[l], W = lv.W[l], st = lv.start[l];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = i == 0 ? xy01.x : i == 1 ? xy01.z : i == 2 ? xy23.x : xy23.z;
        const float y = i == 0 ? xy01.y : i == 1 ? xy01.w : i == 2 ? xy23.y : xy23.w;
        const float a = i == 0 ? aw.x : i == 1 ? aw.y : i == 2 ? aw.z : aw.w;
        int tok[4], h0, w0;
        float lh, lw;
        const bool in = msda_sample_geom_hw(x, y, H, W, st, tok, lh, lw, h0, w0, fma);
        int base = 0, mask = 0;
        if (in) {
          mask = (tok[0] >= 0) | ((tok[1] >= 0) << 1) | ((tok[2] >= 0) << 2) | ((tok[3] >= 0) << 3);
          base = ((st + h0 * W + w0) * M + m) * 32;  // row of (h0, w0), virtual when h0 or w0 is -1
          hmn = min(hmn, h0);
          wmn = min(wmn, w0);
        }
        p_bm[li][i] = base | mask;
        p_h0[li][i] = h0; p_w0[li][i] = w0;
        p_lh[li][i] = lh; p_lw[li][i] = lw; p_a[li][i] = a;
      }
    }
    // the 8 lanes of a warp with the same r decode the same level: combine, then one atomic each
#pragma unroll
    for (int s = 4; s <= 16; s <<= 1) {
      hmn = min(hmn, __shfl_xor_sync(0xffffffffu, hmn, s));
      wmn = min(wmn, __shfl_xor_sync(0xffffffffu, wmn, s));
    }
    if (lane < 4 && l < kL && hmn != INT_MAX) {
      atomicMin(&misc[l], hmn);
      atomicMin(&misc[8 + l], wmn);
    }
  }
  __syncthreads();  // window origins known

  // ---- phase 1b: slots, ranks, records ---------------------------------------------------------
  int p_sr[NLV][4][4];  // (slot << 13) | rank, or -1
#pragma unroll
  for (int li = 0; li < NLV; ++li) {
    const int l = r + 4 * li;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) p_sr[li][i][k] = -1;
    if (l < kL && act) {
      const int hm = misc[l], wm = misc[8 + l];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int bm = p_bm[li][i];
        const bool in = (bm & 15) != 0;
        bool agg = false;
        int dh = 0, dw = 0;
        if (in) {
          dh = p_h0[li][i] - hm;
          dw = p_w0[li][i] - wm;
          agg = (dh + 1 < kAggW) && (dw + 1 < kAggW);
        }
        if (agg) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (bm & (1 << k)) {
              const int slot = l * kAggSlots + (dh + (k >> 1)) * kAggW + dw + (k & 1);
              const int rank = atomicAdd(&hist[slot], 1);
              p_sr[li][i][k] = (slot << 13) | rank;
            }
          }
        }
        // bit 4 = "direct": the point's corners are reduced straight to global memory in phase 2
        rec[ql * Cfg::REC_STRIDE + l * kP + i] =
            make_float4(__int_as_float(bm | ((in && !agg) ? 16 : 0)), p_lh[li][i], p_lw[li][i], p_a[li][i]);
      }
    }
  }
  __syncthreads();  // histogram complete, records visible

  // ---- exclusive scan of the histogram, in place ---------------------------------------------------
  {
    int v[Cfg::SPT], sum = 0;
#pragma unroll
    for (int k = 0; k < Cfg::SPT; ++k) { v[k] = hist[t * Cfg::SPT + k]; sum += v[k]; }
    int inc = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= s) inc += n;
    }
    if (lane == 31) misc[16 + warp] = inc;
    __syncthreads();
    int run = inc - sum;
    for (int w = 0; w < warp; ++w) run += misc[16 + w];
#pragma unroll
    for (int k = 0; k < Cfg::SPT; ++k) { hist[t * Cfg::SPT + k] = run; run += v[k]; }
    if (t == kAggThreads - 1) misc[24] = run;
    __syncthreads();
  }

  // ---- phase 1c: write contributions to their sorted positions -----------------------------------
#pragma unroll
  for (int li = 0; li < NLV; ++li)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lh = p_lh[li][i], lw = p_lw[li][i], a = p_a[li][i];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sr = p_sr[li][i][k];
        if (sr >= 0) {
          const int slot = sr >> 13;
          const float coef = (a * ((k & 2) ? lh : 1.f - lh)) * ((k & 1) ? lw : 1.f - lw);
          contrib[hist[slot] + (sr & 8191)] = make_int2((slot << 6) | ql, __float_as_int(coef));
        }
      }
    }

  // ---- phase 2: gather; grad_sampling_loc, grad_attn_weight; direct reductions -------------------
  const int g = lane / G, j = lane % G;
  const size_t img = (size_t)b * S * M32 + j * C;
  const VT* value_b = value + img;
  float* gvalue_b = grad_value + img;
#pragma unroll 1
  for (int pass = 0; pass < PASSES; ++pass) {
    const int ql2 = pass * QPP + warp * GPW + g;
    const int q2 = qidx[ql2];
    const bool active = q2 >= 0;
    const size_t qm2 = ((size_t)b * Lq + (active ? q2 : 0)) * M + m;
    const unsigned amask = __ballot_sync(0xffffffffu, active);
    if (active) {
      float go[C];
      RT::load_stream(grad_out + qm2 * 32 + j * C, go);
#pragma unroll
      for (int c = 0; c < C; c += 4)
        *reinterpret_cast<float4*>(go_s + ql2 * 32 + j * C + c) = make_float4(go[c], go[c + 1], go[c + 2], go[c + 3]);
      const float4* rq = rec + ql2 * Cfg::REC_STRIDE;
#pragma unroll
      for (int blk = 0; blk < KPG; ++blk) {
        float pgx[G], pgy[G], pga[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int p = blk * G + i;
          pgx[i] = pgy[i] = pga[i] = 0.f;
          if (p < LP) {
            const int l = p / kP;
            const float4 rr = rq[p];
            const int bm = __float_as_int(rr.x);
            const float lh = rr.y, lw = rr.z, a = rr.w;
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float a_hh = a * hh, a_lh = a * lh;
            const ptrdiff_t o0 = (ptrdiff_t)(bm & ~31);
            const ptrdiff_t o2 = o0 + (ptrdiff_t)(lv.W[l] * M32);
            const bool direct = bm & 16;
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (bm & (1 << k)) {
                const ptrdiff_t o = ((k & 2) ? o2 : o0) + ((k & 1) ? M32 : 0);
                float v[C];
                RT::load(value_b + o, v);
                if (direct) {
                  const float tt = ((k & 2) ? a_lh : a_hh) * ((k & 1) ? lw : hw);
#pragma unroll
                  for (int c = 0; c < C; c += 4)
                    red_add_f4(gvalue_b + o + c, tt * go[c], tt * go[c + 1], tt * go[c + 2], tt * go[c + 3]);
                }
                float sdot = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) sdot = fmaf(go[c], v[c], sdot);
                d[k] = sdot;
              }
            }
            pga[i] = hh * (hw * d[0] + lw * d[1]) + lh * (hw * d[2] + lw * d[3]);                  // (cuh:156)
            pgx[i] = (a * (float)lv.W[l]) * (hh * (d[1] - d[0]) + lh * (d[3] - d[2]));            // (cuh:157)
            pgy[i] = (a * (float)lv.H[l]) * (hw * (d[2] - d[0]) + lw * (d[3] - d[1]));            // (cuh:158)
          }
        }
        const float gx = group_reduce_scatter<G>(pgx, j, amask);
        const float gy = group_reduce_scatter<G>(pgy, j, amask);
        const float ga = group_reduce_scatter<G>(pga, j, amask);
        const int p = blk * G + j;
        if (p < LP) {
          st_stream_f2(grad_loc + (qm2 * LP + p) * 2, make_float2(gx, gy));
          st_stream_f1(grad_attw + qm2 * LP + p, ga);
        }
      }
    }
  }
  __syncthreads();  // contributions and grad_out rows are in shared memory

  // ---- phase 3: owner-computes reduce over the sorted contributions ----------------------------------
  {
    constexpr int NG = (kAggThreads / 32) * GPW;
    const int total = misc[24];
    const int chunk = (total + NG - 1) / NG;
    const int gi = warp * GPW + g;
    const int i0 = gi * chunk, i1 = min(total, i0 + chunk);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    int cur = -1;
    auto flush = [&](int slot) {
      const int l = slot / kAggSlots;
      const int rem = slot - l * kAggSlots;
      const int dh = rem / kAggW, dw = rem - dh * kAggW;
      const int tok = lv.start[l] + (misc[l] + dh) * lv.W[l] + misc[8 + l] + dw;
      float* p = gvalue_b + (ptrdiff_t)(tok * M + m) * 32;
#pragma unroll
      for (int c = 0; c < C; c += 4) red_add_f4(p + c, acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    };
    for (int i = i0; i < i1; ++i) {
      const int2 e = contrib[i];
      const int slot = e.x >> 6;
      if (slot != cur) {
        if (cur >= 0) flush(cur);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        cur = slot;
      }
      const float cf = __int_as_float(e.y);
      const float* gp = go_s + (e.x & 63) * 32 + j * C;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 gq = *reinterpret_cast<const float4*>(gp + c);
        acc[c] = fmaf(cf, gq.x, acc[c]); acc[c + 1] = fmaf(cf, gq.y, acc[c + 1]);
        acc[c + 2] = fmaf(cf, gq.z, acc[c + 2]); acc[c + 3] = fmaf(cf, gq.w, acc[c + 3]);
      }
    }
    if (cur >= 0) flush(cur);
  }
}

}  // namespace msda
