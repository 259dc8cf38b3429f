// msda_d32_gv.cuh — grad_value alone, by cell-sorted accumulation (head_dim 32, any value type).
//
// The backward of MSDeformAttn has two independent halves:
//   (1) grad_sampling_loc / grad_attn_weight need the gathered value rows (dot products with grad_out):
//       a gather, bound by L1 line lookups like the forward (msda_bwd_d32_kernel<..., kScatter=false>);
//   (2) grad_value needs NO value rows at all: it is a scatter of weight * grad_out[q] rows, bound by
//       L2's reduction rate when done naively (msda_d32.cuh) and by instruction issue when merged on
//       chip.
// The fused window backward (msda_d32_win.cuh) does both in one block and pays for it in registers
// (128 per thread, 16 warps per SM).  This kernel is half (2) on its own: the same counting sort by
// window cell and the same register accumulation with REDG.E.ADD.F32x4 flushes, but no value window, no
// staging and no dot products — a third of the shared memory and ~64 registers, so three to four
// blocks per SM hide its latencies; half (1) runs as the existing gather kernel.
//
// Gradient formula: models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:125,134,143,152
// (grad_value[corner] += corner weight * attention weight * grad_out).
#pragma once

#include "msda_d32_win.cuh"

namespace msda {

#ifndef MSDA_GV_CELLS
#define MSDA_GV_CELLS 1024
#endif
constexpr int kGvCells = MSDA_GV_CELLS;  // window cells (= value rows) a block can address per head

template <int kL>
struct GvCfg {
  static constexpr int LP = kL * 4;
  static constexpr int NLV = (kL + 3) / 4;
  static constexpr int REC_STRIDE = LP + 1;
  static constexpr int REC_BYTES = kWinTileQ * REC_STRIDE * 16;
  static constexpr int GO_BYTES = kWinTileQ * 32 * 4;
  static constexpr int HIST_N = ((kGvCells + kWinThreads - 1) / kWinThreads) * kWinThreads;
  static constexpr int SPT = HIST_N / kWinThreads;
  static constexpr int HIST_BYTES = (HIST_N + 4) * 4;
  static constexpr int ROWOFF_BYTES = (kGvCells + 4) * 4;
  static constexpr int SORTED_BYTES = ((kWinTileQ * LP * 2 + 15) / 16) * 16 + 16;
  static constexpr int OFF_GO = REC_BYTES;
  static constexpr int OFF_HIST = OFF_GO + GO_BYTES;
  static constexpr int OFF_ROWOFF = OFF_HIST + HIST_BYTES;
  static constexpr int OFF_SORTED = OFF_ROWOFF + ROWOFF_BYTES;
  static constexpr int OFF_MISC = OFF_SORTED + SORTED_BYTES;
  static constexpr int SMEM_BYTES = OFF_MISC + 64 * 4;
  static_assert(kGvCells + 2 < 32768, "two rows are packed in one record word");
};

template <typename VT, int kL, int kM>
__global__ void __launch_bounds__(kWinThreads, 3)
msda_gradvalue_d32_kernel(const VT* __restrict__ grad_out, const float* __restrict__ loc,
                          const float* __restrict__ attw, float* __restrict__ grad_value,
                          const int* __restrict__ order, const int order_len,
                          const __grid_constant__ MsdaLevels lv, const int S, const int M_rt, const int Lq) {
  using Cfg = GvCfg<kL>;
  using RT = RowTraits<VT>;
  constexpr int LP = Cfg::LP, GG = RT::G, GC = RT::C;

  extern __shared__ __align__(128) unsigned char smraw[];
  float4* rec = reinterpret_cast<float4*>(smraw);
  float* go_s = reinterpret_cast<float*>(smraw + Cfg::OFF_GO);
  int* hist = reinterpret_cast<int*>(smraw + Cfg::OFF_HIST);  // counts, then exclusive offsets
  int* rowoff = reinterpret_cast<int*>(smraw + Cfg::OFF_ROWOFF);
  unsigned short* sorted = reinterpret_cast<unsigned short*>(smraw + Cfg::OFF_SORTED);
  int* misc = reinterpret_cast<int*>(smraw + Cfg::OFF_MISC);  // [0,16) warp totals [16] total [32,64) bounding boxes
  int* bb = misc + 32;

  const int M = kM ? kM : M_rt;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int m = blockIdx.x % M, tile = blockIdx.x / M, b = blockIdx.y;
  const int M32 = M * 32;
  const size_t img = (size_t)b * S * M32;

  // ---- decode + bounding boxes: thread = (level slot, query), level warp-uniform -------------------
  const int dql = t & (kWinTileQ - 1), dslot = t / kWinTileQ;
  int dq = -1;
  {
    const int oslot = tile * kWinTileQ + dql;
    if (oslot < order_len) dq = order ? order[oslot] : oslot;
  }
  const size_t dqm = ((size_t)b * Lq + (dq >= 0 ? dq : 0)) * M + m;
  if (t < 32) bb[t] = (t & 8) ? INT_MIN : INT_MAX;  // [0,8) hmin [8,16) hmax [16,24) wmin [24,32) wmax
#pragma unroll
  for (int k = 0; k < Cfg::SPT; ++k) hist[t + k * kWinThreads] = 0;
  // grad_out rows of the tile -> shared memory as fp32
  for (int i = t; i < kWinTileQ * GG; i += kWinThreads) {
    const int gql = i / GG, gj = i % GG;
    const int oslot = tile * kWinTileQ + gql;
    int gq = -1;
    if (oslot < order_len) gq = order ? order[oslot] : oslot;
    float gv[GC];
#pragma unroll
    for (int c = 0; c < GC; ++c) gv[c] = 0.f;
    if (gq >= 0) RT::load_stream(grad_out + (((size_t)b * Lq + gq) * M + m) * 32 + gj * GC, gv);
#pragma unroll
    for (int c = 0; c < GC; c += 4)
      *reinterpret_cast<float4*>(go_s + gql * 32 + gj * GC + c) = make_float4(gv[c], gv[c + 1], gv[c + 2], gv[c + 3]);
  }
  __syncthreads();
  WinPoint pts[Cfg::NLV][4];
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li) {
    const int l = dslot + 4 * li;
    int hmn = INT_MAX, hmx = INT_MIN, wmn = INT_MAX, wmx = INT_MIN;
#pragma unroll
    for (int i = 0; i < 4; ++i) pts[li][i] = WinPoint{0, 0, 0.f, 0.f, 0.f, false};
    if (l < kL) {
      if (dq >= 0) win_decode_level<LP>(loc, attw, dqm, l, lv, pts[li], hmn, hmx, wmn, wmx);
      hmn = __reduce_min_sync(0xffffffffu, hmn); hmx = __reduce_max_sync(0xffffffffu, hmx);
      wmn = __reduce_min_sync(0xffffffffu, wmn); wmx = __reduce_max_sync(0xffffffffu, wmx);
      if (lane == 0 && hmn <= hmx) {
        atomicMin(&bb[l], hmn); atomicMax(&bb[8 + l], hmx);
        atomicMin(&bb[16 + l], wmn); atomicMax(&bb[24 + l], wmx);
      }
    }
  }
  __syncthreads();

  // ---- cells, row offsets, records, histogram -----------------------------------------------------------
  WinAlloc<kL> wa;
  win_allocate<kL, kGvCells>(bb, wa);
  // rowoff[cell row] = element offset of the value / grad_value row inside the image, -1 outside the image
#pragma unroll
  for (int l = kL - 1; l >= 0; --l) {
    if (wa.base[l] < 0) continue;
    const int bw = wa.bw[l], H = lv.H[l], W = lv.W[l];
    for (int rh = warp; rh < wa.bh[l]; rh += kWinThreads / 32) {
      const int h = wa.hm[l] + rh;
      const bool hin = (unsigned)h < (unsigned)H;
      const int row_l = wa.base[l] + rh * bw;
      const int off_l = ((lv.start[l] + h * W + wa.wm[l]) * M + m) * 32;
      for (int rw = lane; rw < bw; rw += 32) {
        const bool inb = hin && (unsigned)(wa.wm[l] + rw) < (unsigned)W;
        rowoff[row_l + rw] = inb ? off_l + rw * M32 : -1;
      }
    }
  }
  if (t < 4) rowoff[kGvCells + t] = -1;
  int rank[Cfg::NLV][4];
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li) {
    const int l = dslot + 4 * li;
    if (l < kL) {
      int base = -1, bw = 0, hm = 0, wm = 0;
#pragma unroll
      for (int ll = 0; ll < kL; ++ll)
        if (ll == l) { base = wa.base[ll]; bw = wa.bw[ll]; hm = wa.hm[ll]; wm = wa.wm[ll]; }
      const int H = lv.H[l], W = lv.W[l], st = lv.start[l];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int code = win_record_code<kGvCells>(pts[li][i], base, bw, hm, wm, H, W, st, m, M);
        rank[li][i] = -1;
        if (base >= 0 && pts[li][i].in) rank[li][i] = atomicAdd(&hist[code & 0xffff], 1);
        rec[dql * Cfg::REC_STRIDE + l * 4 + i] =
            make_float4(__int_as_float(code), pts[li][i].lh, pts[li][i].lw, pts[li][i].in ? pts[li][i].a : 0.f);
      }
    }
  }
  __syncthreads();

  // ---- exclusive scan of the per-cell counts, in place -----------------------------------------------
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
    if (lane == 31) misc[warp] = inc;
    __syncthreads();
    int run = inc - sum;
    for (int w = 0; w < warp; ++w) run += misc[w];
#pragma unroll
    for (int k = 0; k < Cfg::SPT; ++k) { hist[t * Cfg::SPT + k] = run; run += v[k]; }
    if (t == kWinThreads - 1) misc[16] = run;
    __syncthreads();
  }
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li) {
    const int l = dslot + 4 * li;
    if (l < kL) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (rank[li][i] >= 0) {
          const int code = __float_as_int(rec[dql * Cfg::REC_STRIDE + l * 4 + i].x);
          sorted[hist[code & 0xffff] + rank[li][i]] = (unsigned short)(dql * LP + l * 4 + i);
        }
    }
  }
  __syncthreads();

  // ---- sorted pass: 4-lane groups x 8 channels walk contiguous chunks of the cell-sorted samples --------------
  {
    using SL = SortLane<float>;  // only fp32 rows (grad_out copy, grad_value) are touched here
    constexpr int SG = 4, SP = 4, SNG = kWinThreads / SG;
    const int sg = lane >> 2, sj = lane & 3;
    const int oA = SL::chunk_a(sg, sj) * 16, oB = SL::chunk_b(sg, sj) * 16;
    float* gvalue_a = grad_value + img + oA / 4;
    float* gvalue_b = grad_value + img + oB / 4;
    const unsigned char* go_a = reinterpret_cast<const unsigned char*>(go_s) + oA;
    const unsigned char* go_b = reinterpret_cast<const unsigned char*>(go_s) + oB;
    const int total = misc[16];
    const int chunk = (((total + SNG - 1) / SNG) + 3) & ~3;
    const int gi = warp * 8 + sg;
    const int i0 = min(total, gi * chunk), i1 = min(total, i0 + chunk);
    float2 A0[SP], A1[SP], B0[SP], B1[SP];
    const float2 zero2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < SP; ++c) A0[c] = A1[c] = B0[c] = B1[c] = zero2;
    int cur0 = -2, cur1 = -2;
    auto flush = [&](const int row, const float2 (&acc)[SP]) {
      const int off = rowoff[row];
      if (off >= 0) {
        red_add_f4(gvalue_a + off, acc[0].x, acc[0].y, acc[1].x, acc[1].y);
        red_add_f4(gvalue_b + off, acc[2].x, acc[2].y, acc[3].x, acc[3].y);
      }
    };
    auto rec_slot = [&](const int sid) { const int sq = sid / LP; return sq * Cfg::REC_STRIDE + (sid - sq * LP); };
    float4 rnext = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 < i1) rnext = rec[rec_slot(sorted[i0])];
    for (int ib = i0; ib < i1; ib += 4) {
      const uint2 packed = *reinterpret_cast<const uint2*>(sorted + ib);  // 4 sample ids (ib is a multiple of 4)
      const int nb = min(4, i1 - ib);
      int nsid = 0;
      if (ib + 4 < i1) nsid = sorted[ib + 4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (u < nb) {  // group-uniform
          const int sid = (int)(((u < 2 ? packed.x : packed.y) >> ((u & 1) * 16)) & 0xffffu);
          const int sq = sid / LP;
          const float4 r = rnext;
          const float4 ga4 = *reinterpret_cast<const float4*>(go_a + sq * 128);
          const float4 gb4 = *reinterpret_cast<const float4*>(go_b + sq * 128);
          {
            const int sidn = u == 3 ? nsid
                                    : (int)((((u + 1) < 2 ? packed.x : packed.y) >> (((u + 1) & 1) * 16)) & 0xffffu);
            if (u + 1 < nb || (u == 3 && ib + 4 < i1)) rnext = rec[rec_slot(sidn)];
          }
          const int code = __float_as_int(r.x);
          const int row0 = code & 0xffff, row1 = code >> 16;
          WIN_CHECK(sid < kWinTileQ * LP && row0 >= cur0 && row1 + 1 < kGvCells && row1 > row0);
          if (row0 != cur0) {
            const bool adj = (row0 == cur0 + 1);
            if (cur0 >= 0) {
              flush(cur0, A0);
              flush(cur1, B0);
              if (!adj) {
                flush(cur0 + 1, A1);
                flush(cur1 + 1, B1);
              }
            }
#pragma unroll
            for (int c = 0; c < SP; ++c) {
              A0[c] = adj ? A1[c] : zero2;
              B0[c] = adj ? B1[c] : zero2;
              A1[c] = zero2;
              B1[c] = zero2;
            }
            cur0 = row0; cur1 = row1;
          }
          const float lh = r.y, lw = r.z, a = r.w;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float2 go[SP] = {make_float2(ga4.x, ga4.y), make_float2(ga4.z, ga4.w), make_float2(gb4.x, gb4.y),
                                 make_float2(gb4.z, gb4.w)};
          const float w00 = (hh * hw) * a, w01 = (hh * lw) * a, w10 = (lh * hw) * a, w11 = (lh * lw) * a;
          const float2 w00p = make_float2(w00, w00), w01p = make_float2(w01, w01), w10p = make_float2(w10, w10),
                       w11p = make_float2(w11, w11);
#pragma unroll
          for (int c = 0; c < SP; ++c) {
            A0[c] = ffma2(w00p, go[c], A0[c]); A1[c] = ffma2(w01p, go[c], A1[c]);
            B0[c] = ffma2(w10p, go[c], B0[c]); B1[c] = ffma2(w11p, go[c], B1[c]);
          }
        }
      }
    }
    if (cur0 >= 0) {
      flush(cur0, A0);
      flush(cur1, B0);
      flush(cur0 + 1, A1);
      flush(cur1 + 1, B1);
    }
  }

  // ---- direct pass: levels whose bounding box has more cells than the block can address -------------------
  bool all_cells = true;
#pragma unroll
  for (int l = 0; l < kL; ++l) all_cells = all_cells && wa.base[l] >= 0;
  if (!all_cells) {
    const int g = lane >> 3, j = lane & 7;  // 8 lanes x float4 per fp32 row
    float* gvalue_j = grad_value + img + j * 4;
    for (int ql = warp * 4 + g; ql < kWinTileQ; ql += kWinThreads / 8) {
      const float4 go = *reinterpret_cast<const float4*>(go_s + ql * 32 + j * 4);
#pragma unroll
      for (int l = 0; l < kL; ++l) {
        if (wa.base[l] >= 0) continue;  // block-uniform
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 r = rec[ql * Cfg::REC_STRIDE + l * 4 + i];
          const int code = __float_as_int(r.x);
          const float lh = r.y, lw = r.z, a = r.w;
          const float a_hh = a * (1.f - lh), a_lh = a * lh, hw = 1.f - lw;
          const ptrdiff_t o0 = (ptrdiff_t)(code & ~31);
          const ptrdiff_t o2 = o0 + (ptrdiff_t)(lv.W[l] * M32);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (code & (1 << k)) {
              const ptrdiff_t o = ((k & 2) ? o2 : o0) + ((k & 1) ? M32 : 0);
              const float tt = ((k & 2) ? a_lh : a_hh) * ((k & 1) ? lw : hw);
              red_add_f4(gvalue_j + o, tt * go.x, tt * go.y, tt * go.z, tt * go.w);
            }
          }
        }
      }
    }
  }
}

}  // namespace msda
