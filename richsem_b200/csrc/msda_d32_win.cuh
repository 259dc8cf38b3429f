// msda_d32_win.cuh — "window" kernels for head_dim 32: the value rows a block of queries samples
// are staged ONCE in shared memory; the forward gathers from there, the backward walks the block's
// samples sorted by window cell so that value rows AND grad_value partial sums live in registers.
//
// Why (measured on B200, scratch/gather_bw.cu, scratch/tma_stage.cu, scratch/smem_atomics.cu):
//   * a gather of 128-byte rows out of L1 (LDG.128, 4 rows per warp instruction, all hits) runs at
//     1.2 SM-cycles per row with all of L1 available and 1.9 when shared memory takes most of it (the
//     tiled kernels of msda_d32.cuh sit at 1.8), while the same gather out of shared memory (LDS.128)
//     runs at 1.04 cycles per row, the 128 B/clk limit of the data pipe;
//   * L2 retires scattered fp32 row reductions at 5.8 SM-cycles per row per SM (6.4 TB/s chip-wide):
//     22.75 M of them per bs=2 encoder layer is the 455 us floor of the plain backward; shared-memory
//     float atomics are a CAS loop (14.7 cycles per row; 5.1 with a 128-bit CAS), so merging must be
//     "owner computes", and native int ATOMS.ADD (0.1 cycles per lane) makes a counting sort cheap;
//   * encoder self-attention is spatially local: the 64 queries of an 8x8-pixel patch put their
//     64*L*P*4 = 4096 corner reads per head on ~330 distinct rows (scratch/bbox_stats.py).
//
// Block = one head x a tile of 64 queries (host-provided patch order for encoder self-attention).
//   front end  thread (level, query) decodes 4 sampling points (bit-exact geometry of msda_common.cuh);
//              REDUX + shared atomics give each level's bounding box of (h0, w0).  Levels whose box
//              [hmin, hmax+1] x [wmin, wmax+1] fits what is left of the row pool get a window (coarsest
//              level first: smallest boxes); warps copy whole window lines with 16-byte cp.async,
//              zero-filling rows outside the image, so windowed samples need no corner predicates.
//              Levels that do not fit stay "direct": their samples gather from global memory.
//   forward    lane groups (8 lanes x float4 for fp32 rows, 4 lanes x 8 bf16) walk the records of their
//              queries: one broadcast LDS.128 per point, four row reads, 16 FFMA.
//   backward   produce: the windowed samples are counting-sorted by cell (= pool row of corner (h0,w0)).
//              consume: each 4-lane group (8 channels per lane) walks a contiguous chunk of the sorted list
//              holding the current cell's four value rows and four grad_value accumulators in registers:
//              per sample two LDS.128 of grad_out, 16 FFMA2 of dot products, 16 FFMA2 of accumulation; a
//              cell change flushes two (adjacent cell: the other two slide over) or four accumulators with
//              REDG.ADD.F32x4 — 4.4x fewer L2 reductions than one per corner.  grad_sampling_loc /
//              grad_attn_weight are parked in the record slots and written out coalesced at the end.
//              Variants of the same two halves: deterministic (canonical order inside the block, 64-bit
//              fixed-point accumulation across blocks) and persistent / warp-specialised (opt-in).
//
// The arithmetic restates models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299 (forward),
// :87-159 (gradients); nothing of that file's thread mapping or reductions is used.
#pragma once

#include <climits>

#include "msda_d32.cuh"

namespace msda {

// Queries per block of the window kernels; 4 threads per query (one per level slot).  The host's padded
// 8x8-patch order serves 64 (one patch) as well as 32 (the upper / lower 8x4 half of a patch).
#ifndef MSDA_WIN_TILE
#define MSDA_WIN_TILE 64
#endif
constexpr int kWinTileQ = MSDA_WIN_TILE;
constexpr int kWinThreads = 4 * kWinTileQ;
static_assert(kWinTileQ == 32 || kWinTileQ == 64 || kWinTileQ == 128, "window kernels: 32, 64 or 128 queries per block");
// Rows of the window pool.  Forward: 448 rows = 75 KB per block with the records, three blocks per SM.
// Backward: 448 rows = 89 KB with the sort structures, two blocks per SM (128 registers per thread);
// measured per bs=2 encoder layer: 256 rows 0.464 ms, 320 0.442, 384 0.425, 448 0.414, 592 0.434.
#ifndef MSDA_WIN_POOL_FWD
#define MSDA_WIN_POOL_FWD 448
#endif
#ifndef MSDA_WIN_POOL_BWD
#define MSDA_WIN_POOL_BWD 448
#endif
constexpr int kWinPoolFwd = MSDA_WIN_POOL_FWD;
constexpr int kWinPoolBwd = MSDA_WIN_POOL_BWD;

#ifdef MSDA_WIN_TIMING
// Phase timing (debug builds only): per-phase SM-clock sums over all blocks, read by msda_debug_win_timing().
__device__ unsigned long long g_win_timing[16];
#define WIN_T(k, t0)                                                                          \
  do {                                                                                        \
    if ((threadIdx.x % kWinThreads) == 0) {                                                   \
      const long long now_ = clock64();                                                       \
      atomicAdd(&g_win_timing[k], (unsigned long long)(now_ - (t0)));                         \
      (t0) = now_;                                                                            \
    }                                                                                         \
  } while (0)
#else
#define WIN_T(k, t0) do { } while (0)
#endif

// Index checks for debug builds (MSDA_NVCC_EXTRA=-DMSDA_WIN_CHECKS): compute-sanitizer is not available on
// the GPU pool, so the shared-memory indices of these kernels are asserted by hand; a failed check traps.
#ifdef MSDA_WIN_CHECKS
#define WIN_CHECK(cond) do { if (!(cond)) asm volatile("trap;"); } while (0)
#else
#define WIN_CHECK(cond) do { } while (0)
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// 16-byte global->shared copy that bypasses L1 and registers; src_bytes = 0 writes zeros.
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// Pulls the line holding p into L2 (no register, no L1): used one wave of blocks ahead of the demand load.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// How many tiles ahead a block prefetches sampling locations / weights / grad_out (~ one wave of blocks:
// 148 SMs x 2..3 blocks / 8 heads).
constexpr int kWinPrefetchTiles = 48 * 64 / kWinTileQ;

__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Shared-memory row reads of the window (VT rows of 32 channels).
template <typename VT>
struct WinRow;
template <>
struct WinRow<float> {
  static constexpr int ROWB = 128;
  static __device__ __forceinline__ void lds(const unsigned char* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct WinRow<__nv_bfloat16> {
  static constexpr int ROWB = 64;
  static __device__ __forceinline__ void lds(const unsigned char* p, float (&v)[8]) {
    RowTraits<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(p), v);
  }
};

template <typename VT, int kL, int kWinPool>
struct WinCfg {
  static constexpr int LP = kL * 4;
  static constexpr int NLV = (kL + 3) / 4;            // levels decoded per thread
  static constexpr int REC_STRIDE = LP + 1;           // float4 per query (+1: bank skew)
  static constexpr int ROWB = WinRow<VT>::ROWB;
  static constexpr int POOL_BYTES = (kWinPool + 2) * ROWB;  // + two all-zero rows for skipped samples
  static constexpr int REC_BYTES = kWinTileQ * REC_STRIDE * 16;
  static constexpr int BB_BYTES = 32 * 4;
  static constexpr int FWD_SMEM = POOL_BYTES + REC_BYTES + BB_BYTES;
  // backward extras
  static constexpr int HIST_N = ((kWinPool + kWinThreads - 1) / kWinThreads) * kWinThreads;  // padded for the scan
  static constexpr int SPT = HIST_N / kWinThreads;
  static constexpr int GO_BYTES = kWinTileQ * 32 * 4;            // grad_out rows of the tile, fp32
  static constexpr int HIST_BYTES = (HIST_N + 4) * 4;         // counts -> offsets (+ total)
  static constexpr int ROWOFF_BYTES = (kWinPool + 2) * 4;
  static constexpr int SORTED_BYTES = ((kWinTileQ * LP * 2 + 15) / 16) * 16 + 16;
  static constexpr int OFF_GO = POOL_BYTES + REC_BYTES + BB_BYTES;
  static constexpr int OFF_HIST = OFF_GO + GO_BYTES;
  static constexpr int OFF_ROWOFF = OFF_HIST + HIST_BYTES;
  static constexpr int OFF_SORTED = OFF_ROWOFF + ROWOFF_BYTES;
  static constexpr int INFLAG_BYTES = ((NLV * 4 * kWinTileQ + 127) / 128) * 128;
  static constexpr int BWD_SMEM = OFF_SORTED + SORTED_BYTES + 192 + INFLAG_BYTES + kWinTileQ * 8;  // + softmax stats
  // deterministic mode: per-warp, per-cell sample counts (16-bit) that make the sort ranks scheduling-independent
  static constexpr int WCNT_BYTES = (kWinThreads / 32) * HIST_N * 2;
  static constexpr int BWD_DET_SMEM = ((BWD_SMEM + 15) / 16) * 16 + WCNT_BYTES;
  static constexpr int BWD_SET_BYTES = ((BWD_SMEM + 127) / 128) * 128;   // one buffer set of the warp-specialised kernel
  static constexpr int BWD_WS_SMEM = 2 * BWD_SET_BYTES + 64;
  static_assert(kL <= 8, "per-level state is kept in 8-entry arrays");
  static_assert(kWinPool + 2 < 32768, "two pool rows are packed in one record word");
  static_assert(kWinTileQ * LP < 65536, "sample ids are stored as 16-bit");
};

// Per-level window decision, computed identically by every thread from the block's bounding boxes.
template <int kL>
struct WinAlloc {
  int base[kL];        // first pool row of the level's window, -1: level is gathered from global memory
  int bw[kL], bh[kL];  // window size in rows
  int hm[kL], wm[kL];  // window origin (pixel coordinates, may be -1)
};

template <int kL, int kWinPool>
__device__ __forceinline__ void win_allocate(const int* bb, WinAlloc<kL>& wa) {
  // bb: [0,8) hmin  [8,16) hmax  [16,24) wmin  [24,32) wmax
  int used = 0;
#pragma unroll
  for (int l = kL - 1; l >= 0; --l) {
    const int hm = bb[l], hM = bb[8 + l], wm = bb[16 + l], wM = bb[24 + l];
    wa.base[l] = -1;
    wa.bw[l] = 2; wa.bh[l] = 2; wa.hm[l] = 0; wa.wm[l] = 0;
    if (hm <= hM) {
      const unsigned bh = (unsigned)(hM - hm) + 2u, bw = (unsigned)(wM - wm) + 2u;
      if (bh <= (unsigned)kWinPool && bw <= (unsigned)kWinPool && used + (int)(bh * bw) <= kWinPool) {
        wa.base[l] = used;
        wa.bw[l] = (int)bw; wa.bh[l] = (int)bh; wa.hm[l] = hm; wa.wm[l] = wm;
        used += (int)(bh * bw);
      }
    }
  }
}

// How the threads that run a phase together synchronise: the whole block, or one 'kWinThreads'-wide group
// of a warp-specialised block (named barrier kId).
struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int kId>
struct GroupSync {
  __device__ __forceinline__ void operator()() const {
    asm volatile("bar.sync %0, %1;" ::"n"(kId), "n"(kWinThreads) : "memory");
  }
};

// Stages the windows of all allocated levels: warp w copies window lines w, w+8, ... of each level.
// rowoff (backward only): per pool row, the element offset of the row inside the image's value block,
// or -1 for rows outside the image.
template <typename VT, int kL, bool kRowOff, int kWinPoolCheck>
__device__ __forceinline__ void win_stage(const WinAlloc<kL>& wa, const MsdaLevels& lv, const VT* value_img,
                                          const int m, const int M, unsigned char* pool, int* rowoff, const int t) {
  constexpr int ROWB = WinRow<VT>::ROWB, G = ROWB / 16, EPL = 16 / (int)sizeof(VT);  // elements per 16 B
  const int warp = t >> 5, lane = t & 31;
  const int rw0 = lane / G, jj = lane % G;
  const unsigned pool_s = smem_u32(pool);
#pragma unroll
  for (int l = kL - 1; l >= 0; --l) {
    if (wa.base[l] < 0) continue;
    const int bw = wa.bw[l], H = lv.H[l], W = lv.W[l];
    for (int rh = warp; rh < wa.bh[l]; rh += kWinThreads / 32) {
      const int h = wa.hm[l] + rh;
      const bool hin = (unsigned)h < (unsigned)H;
      const int row_l = wa.base[l] + rh * bw;                            // pool row of the line's first row
      const int off_l = ((lv.start[l] + h * W + wa.wm[l]) * M + m) * 32;  // its element offset (virtual if outside)
      for (int rw = rw0; rw < bw; rw += 32 / G) {
        const bool inb = hin && (unsigned)(wa.wm[l] + rw) < (unsigned)W;
        const int off = off_l + rw * (M * 32);
        WIN_CHECK(row_l + rw >= 0 && row_l + rw < kWinPoolCheck);
        cp_async16(pool_s + (unsigned)((row_l + rw) * ROWB + jj * 16), value_img + (inb ? off : 0) + jj * EPL,
                   inb ? 16 : 0);
        if (kRowOff && jj == 0) rowoff[row_l + rw] = inb ? off : -1;
      }
    }
  }
}

// One decoded sampling point held in registers between the phases.
struct WinPoint {
  int h0, w0;
  float lh, lw, a;
  bool in;
};

// Record word of a point.
//   level served from the window: pool row of corner (h0,w0) | pool row of corner (h1,w0) << 16; the
//     (.,w1) corners are the next rows.  A skipped sample points at the pool's two all-zero rows (and
//     carries weight 0), so the gather loop needs no branch.
//   level gathered from global memory: element offset of corner (h0,w0)'s row | 4-bit corner mask
//     (as msda_d32.cuh); skipped -> 0.
template <int kWinPool>
__device__ __forceinline__ int win_record_code(const WinPoint& pt, const int base, const int bw, const int hm,
                                               const int wm, const int H, const int W, const int start,
                                               const int m, const int M) {
  if (base >= 0) {
    if (!pt.in) return kWinPool | (kWinPool << 16);
    const int row0 = base + (pt.h0 - hm) * bw + (pt.w0 - wm);
    return row0 | ((row0 + bw) << 16);
  }
  if (!pt.in) return 0;
  const bool h0ok = pt.h0 >= 0, w0ok = pt.w0 >= 0, h1ok = pt.h0 + 1 <= H - 1, w1ok = pt.w0 + 1 <= W - 1;
  const int mask = (h0ok && w0ok) | ((h0ok && w1ok) << 1) | ((h1ok && w0ok) << 2) | ((h1ok && w1ok) << 3);
  return mask ? (((start + pt.h0 * W + pt.w0) * M + m) * 32) | mask : 0;
}

// Decodes the 4 points of level l of (query, head) qm and folds them into the thread's bounding box.
// Softmax statistics of the L*P logits of one (query, head): max and sum of exp(x - max).  Every thread that
// needs them reads the LP contiguous floats itself (the 4 level threads of a query sit in different warps;
// re-reading 64 bytes from L2 is cheaper than a block barrier).
template <int kLP>
__device__ __forceinline__ float2 win_softmax_stats(const float* __restrict__ logits) {
  float4 v[kLP / 4];
#pragma unroll
  for (int i = 0; i < kLP / 4; ++i) v[i] = ld_stream_f4(logits + 4 * i);
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kLP / 4; ++i) mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLP / 4; ++i) sum += __expf(v[i].x - mx) + __expf(v[i].y - mx) + __expf(v[i].z - mx) + __expf(v[i].w - mx);
  return make_float2(mx, sum);
}

template <int kLP>
__device__ __forceinline__ void win_decode_loaded(float4 xy01, float4 xy23, float4 aw, const float* __restrict__ attw,
                                                  const size_t qm, const int l, const MsdaLevels& lv,
                                                  WinPoint (&pt)[4], int& hmn, int& hmx, int& wmn, int& wmx,
                                                  const MsdaFused fz, const size_t bq, float2* stats);

template <int kLP>
__device__ __forceinline__ void win_decode_level(const float* __restrict__ loc, const float* __restrict__ attw,
                                                 const size_t qm, const int l, const MsdaLevels& lv,
                                                 WinPoint (&pt)[4], int& hmn, int& hmx, int& wmn, int& wmx,
                                                 const MsdaFused fz = MsdaFused{nullptr, 0}, const size_t bq = 0,
                                                 float2* stats = nullptr) {
  const float* lp = loc + (qm * kLP + l * 4) * 2;
  const float4 xy01 = ld_stream_f4(lp), xy23 = ld_stream_f4(lp + 4);
  const float4 aw = ld_stream_f4(attw + qm * kLP + l * 4);
  win_decode_loaded<kLP>(xy01, xy23, aw, attw, qm, l, lv, pt, hmn, hmx, wmn, wmx, fz, bq, stats);
}

// The same, with the level's raw 48 bytes already in registers (the front end issues those loads before its
// first barrier).
template <int kLP>
__device__ __forceinline__ void win_decode_loaded(float4 xy01, float4 xy23, float4 aw, const float* __restrict__ attw,
                                                  const size_t qm, const int l, const MsdaLevels& lv,
                                                  WinPoint (&pt)[4], int& hmn, int& hmx, int& wmn, int& wmx,
                                                  const MsdaFused fz, const size_t bq, float2* stats) {
  constexpr int LP = kLP, num_levels = kLP / 4;
  const int H = lv.H[l], W = lv.W[l];
  const bool fma = lv.coord_fma != 0;
  if (fz.ref_dim) {  // fused prologue: raw offsets / logits -> locations / weights
    const float2 st = stats ? *stats : win_softmax_stats<kLP>(attw + qm * LP);  // published by the front end, if any
    aw = make_float4(__expf(aw.x - st.x) / st.y, __expf(aw.y - st.x) / st.y, __expf(aw.z - st.x) / st.y,
                     __expf(aw.w - st.x) / st.y);
    const float2 p0 = msda_fused_location(fz, bq, num_levels, l, 4, H, W, make_float2(xy01.x, xy01.y));
    const float2 p1 = msda_fused_location(fz, bq, num_levels, l, 4, H, W, make_float2(xy01.z, xy01.w));
    const float2 p2 = msda_fused_location(fz, bq, num_levels, l, 4, H, W, make_float2(xy23.x, xy23.y));
    const float2 p3 = msda_fused_location(fz, bq, num_levels, l, 4, H, W, make_float2(xy23.z, xy23.w));
    xy01 = make_float4(p0.x, p0.y, p1.x, p1.y);
    xy23 = make_float4(p2.x, p2.y, p3.x, p3.y);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x = i == 0 ? xy01.x : i == 1 ? xy01.z : i == 2 ? xy23.x : xy23.z;
    const float y = i == 0 ? xy01.y : i == 1 ? xy01.w : i == 2 ? xy23.y : xy23.w;
    pt[i].a = i == 0 ? aw.x : i == 1 ? aw.y : i == 2 ? aw.z : aw.w;
    int tok[4];
    pt[i].in = msda_sample_geom_hw(x, y, H, W, 0, tok, pt[i].lh, pt[i].lw, pt[i].h0, pt[i].w0, fma);
    if (pt[i].in) {
      hmn = min(hmn, pt[i].h0); hmx = max(hmx, pt[i].h0);
      wmn = min(wmn, pt[i].w0); wmx = max(wmx, pt[i].w0);
    }
  }
}

// Front end shared by the forward and backward window kernels: decode, bounding boxes, window
// allocation, staging (cp.async left in flight), records.  Thread t decodes level (t / 64) [+4] of
// query t % 64, so the level is warp-uniform.  kBwd additionally fills rowoff and counts the windowed
// samples per cell (hist; rank[][] = the sample's arrival order inside its cell).
template <typename VT, int kL, int kWinPool, bool kBwd, bool kDetRank, class Sync>
__device__ __forceinline__ void win_front_end(const int t, const Sync sync, unsigned short* wcnt, const VT* __restrict__ value_img, const float* __restrict__ loc,
                                              const float* __restrict__ attw, const int q, const size_t qm,
                                              const MsdaFused fz, const size_t bq, float2* stats,
                                              const int qpf, const size_t qm_pf,
                                              const int m, const int M, const MsdaLevels& lv,
                                              unsigned char* pool, float4* rec, int* bb, int* rowoff, int* hist,
                                              WinAlloc<kL>& wa, WinPoint (&pts)[WinCfg<VT, kL, kWinPool>::NLV][4],
                                              int (&rank)[WinCfg<VT, kL, kWinPool>::NLV][4], long long& tph) {
  using Cfg = WinCfg<VT, kL, kWinPool>;
  const int lane = t & 31;
  const int ql = t & (kWinTileQ - 1), slot = t / kWinTileQ;
  // the decode's global loads go out first: their latency then covers the initialisation and the first barrier
  float4 rxy01[Cfg::NLV], rxy23[Cfg::NLV], raw[Cfg::NLV];
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li) {
    const int l = slot + 4 * li;
    rxy01[li] = rxy23[li] = raw[li] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < kL && q >= 0) {
      const float* lp = loc + (qm * Cfg::LP + l * 4) * 2;
      rxy01[li] = ld_stream_f4(lp);
      rxy23[li] = ld_stream_f4(lp + 4);
      raw[li] = ld_stream_f4(attw + qm * Cfg::LP + l * 4);
    }
  }
  // fused prologue: the level-slot-0 thread of each query takes the softmax statistics of its L*P logits and
  // publishes them through shared memory (the consume half needs them there anyway); the barrier below orders it
  if (fz.ref_dim && slot == 0 && q >= 0 && stats != nullptr) *stats = win_softmax_stats<Cfg::LP>(attw + qm * Cfg::LP);
  if (t < 32) bb[t] = (t & 8) ? INT_MIN : INT_MAX;  // [0,8) hmin [8,16) hmax [16,24) wmin [24,32) wmax
  if (t >= 32 && t < 32 + 2 * Cfg::ROWB / 16)
    reinterpret_cast<uint4*>(pool + kWinPool * Cfg::ROWB)[t - 32] = make_uint4(0u, 0u, 0u, 0u);
  if (kBwd) {
#pragma unroll
    for (int k = 0; k < Cfg::SPT; ++k) hist[t + k * kWinThreads] = 0;
  }
  if (kDetRank) {
    for (int i = t; i < Cfg::WCNT_BYTES / 16; i += kWinThreads) reinterpret_cast<uint4*>(wcnt)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  sync();
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li) {
    const int l = slot + 4 * li;  // warp-uniform
    int hmn = INT_MAX, hmx = INT_MIN, wmn = INT_MAX, wmx = INT_MIN;
#pragma unroll
    for (int i = 0; i < 4; ++i) pts[li][i] = WinPoint{0, 0, 0.f, 0.f, 0.f, false};
    if (l < kL) {
      if (q >= 0)
        win_decode_loaded<Cfg::LP>(rxy01[li], rxy23[li], raw[li], attw, qm, l, lv, pts[li], hmn, hmx, wmn, wmx, fz, bq, stats);
      if (qpf >= 0) {  // a later block's inputs: HBM -> L2 now, so that its decode sees L2 latency
        prefetch_l2(loc + (qm_pf * Cfg::LP + l * 4) * 2);
        prefetch_l2(attw + qm_pf * Cfg::LP + l * 4);
      }
      hmn = __reduce_min_sync(0xffffffffu, hmn); hmx = __reduce_max_sync(0xffffffffu, hmx);
      wmn = __reduce_min_sync(0xffffffffu, wmn); wmx = __reduce_max_sync(0xffffffffu, wmx);
      if (lane == 0 && hmn <= hmx) {
        atomicMin(&bb[l], hmn); atomicMax(&bb[8 + l], hmx);
        atomicMin(&bb[16 + l], wmn); atomicMax(&bb[24 + l], wmx);
      }
    }
  }
  sync();
  if (!kBwd) WIN_T(0, tph);  // decode + bounding boxes
  win_allocate<kL, kWinPool>(bb, wa);
  win_stage<VT, kL, kBwd, kWinPool>(wa, lv, value_img, m, M, pool, rowoff, t);
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li) {
    const int l = slot + 4 * li;
    if (l < kL) {
      int base = -1, bw = 0, hm = 0, wm = 0;
#pragma unroll
      for (int ll = 0; ll < kL; ++ll)
        if (ll == l) { base = wa.base[ll]; bw = wa.bw[ll]; hm = wa.hm[ll]; wm = wa.wm[ll]; }
      const int H = lv.H[l], W = lv.W[l], st = lv.start[l];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int code = win_record_code<kWinPool>(pts[li][i], base, bw, hm, wm, H, W, st, m, M);
        rank[li][i] = -1;
        WIN_CHECK(base < 0 || ((code & 0xffff) + 1 <= kWinPool + 1 && (code >> 16) + 1 <= kWinPool + 1));
        WIN_CHECK(base < 0 || !pts[li][i].in || ((code >> 16) + 1 < kWinPool && (code & 0xffff) >= base));
        if (kBwd && !kDetRank && base >= 0 && pts[li][i].in) rank[li][i] = atomicAdd(&hist[code & 0xffff], 1);
        if (kDetRank) {
          // scheduling-independent rank: lanes of a warp that hit the same cell are ranked in lane order on
          // top of what the warp counted for that cell in earlier rounds; the warps' counts are turned into
          // exclusive bases after the block barrier (win_bwd_produce)
          const bool part = base >= 0 && pts[li][i].in;
          const int cell = code & 0xffff;
          const unsigned peers = __match_any_sync(0xffffffffu, part ? (unsigned)cell : (0xffff0000u | (unsigned)lane));
          const int leader = __ffs(peers) - 1;
          unsigned short* cnt = wcnt + (t >> 5) * Cfg::HIST_N + cell;
          int prev = 0;
          if (part && lane == leader) { prev = *cnt; *cnt = (unsigned short)(prev + __popc(peers)); }
          prev = __shfl_sync(0xffffffffu, prev, leader);
          if (part) rank[li][i] = prev + __popc(peers & ((1u << lane) - 1u));
          __syncwarp();
        }
        rec[ql * Cfg::REC_STRIDE + l * 4 + i] =
            make_float4(__int_as_float(code), pts[li][i].lh, pts[li][i].lw, pts[li][i].in ? pts[li][i].a : 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
#ifndef MSDA_WIN_FWD_MINBLOCKS
#define MSDA_WIN_FWD_MINBLOCKS 2
#endif
template <typename VT, int kL, int kM>
__global__ void __launch_bounds__(kWinThreads, MSDA_WIN_FWD_MINBLOCKS)
msda_fwd_d32_win_kernel(const VT* __restrict__ value, const float* __restrict__ loc,
                        const float* __restrict__ attw, VT* __restrict__ out,
                        const int* __restrict__ order, const int order_len,
                        const __grid_constant__ MsdaLevels lv, const int S, const int M_rt, const int Lq) {
  constexpr int kWinPool = kWinPoolFwd;
  using Cfg = WinCfg<VT, kL, kWinPool>;
  using RT = RowTraits<VT>;
  using WR = WinRow<VT>;
  constexpr int LP = Cfg::LP, G = RT::G, C = RT::C, GPW = 32 / G, ROWB = Cfg::ROWB;
  constexpr int QPP = (kWinThreads / 32) * GPW, PASSES = kWinTileQ / QPP;
  static_assert(kWinTileQ * 4 == kWinThreads, "decode maps 4 threads to a query");

  extern __shared__ __align__(128) unsigned char smraw[];
  unsigned char* pool = smraw;
  float4* rec = reinterpret_cast<float4*>(smraw + Cfg::POOL_BYTES);
  int* bb = reinterpret_cast<int*>(smraw + Cfg::POOL_BYTES + Cfg::REC_BYTES);

  const int M = kM ? kM : M_rt;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int m = blockIdx.x % M, tile = blockIdx.x / M, b = blockIdx.y;
  const int M32 = M * 32;
  const VT* value_img = value + (size_t)b * S * M32;

  // ---- front end: decode, windows, records ---------------------------------------------------
  long long tphase = clock64();
  (void)tphase;
  WinAlloc<kL> wa;
  {
    int q = -1;
    const int oslot = tile * kWinTileQ + (t & (kWinTileQ - 1));
    if (oslot < order_len) q = order ? order[oslot] : oslot;
    const size_t qm = ((size_t)b * Lq + (q >= 0 ? q : 0)) * M + m;
    int qpf = -1;
    const int pslot = oslot + kWinPrefetchTiles * kWinTileQ;
    if (pslot < order_len) qpf = order ? order[pslot] : pslot;
    const size_t qm_pf = ((size_t)b * Lq + (qpf >= 0 ? qpf : 0)) * M + m;
    WinPoint pts[Cfg::NLV][4];
    int rank[Cfg::NLV][4];
    win_front_end<VT, kL, kWinPool, false, false>(t, BlockSync{}, nullptr, value_img, loc, attw, q, qm, MsdaFused{nullptr, 0}, 0, nullptr, qpf, qm_pf, m, M, lv, pool, rec, bb, nullptr, nullptr, wa, pts, rank, tphase);
  }
  WIN_T(1, tphase);  // allocation, staging issue, records
  cp_async_wait_all();
  __syncthreads();
  WIN_T(2, tphase);  // wait for the windows

  // ---- gather --------------------------------------------------------------------------------
  const int g = lane / G, j = lane % G;
  const unsigned char* pool_j = pool + j * 16;
  const VT* value_j = value_img + j * C;
  bool all_win = true;
#pragma unroll
  for (int l = 0; l < kL; ++l) all_win = all_win && wa.base[l] >= 0;
#pragma unroll 1
  for (int pass = 0; pass < PASSES; ++pass) {
    const int ql = pass * QPP + warp * GPW + g;
    const int oslot = tile * kWinTileQ + ql;
    int q = -1;
    if (oslot < order_len) q = order ? order[oslot] : oslot;
    if (q < 0) continue;
    const size_t qm = ((size_t)b * Lq + q) * M + m;
    const float4* rq = rec + ql * Cfg::REC_STRIDE;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    // one sampling point served from the window: no predicates, no branches
    auto from_window = [&](const float4 r) {
      const int code = __float_as_int(r.x);
      const float lh = r.y, lw = r.z, a = r.w;
      const float a_hh = a * (1.f - lh), a_lh = a * lh, hw = 1.f - lw;
      const unsigned char* p0 = pool_j + (code & 0xffff) * ROWB;
      const unsigned char* p2 = pool_j + (code >> 16) * ROWB;
      float v00[C], v01[C], v10[C], v11[C];
      WR::lds(p0, v00); WR::lds(p0 + ROWB, v01); WR::lds(p2, v10); WR::lds(p2 + ROWB, v11);
      const float w00 = a_hh * hw, w01 = a_hh * lw, w10 = a_lh * hw, w11 = a_lh * lw;
#pragma unroll
      for (int c = 0; c < C; ++c)
        acc[c] = fmaf(w11, v11[c], fmaf(w10, v10[c], fmaf(w01, v01[c], fmaf(w00, v00[c], acc[c]))));
    };
    if (all_win) {
#pragma unroll
      for (int p = 0; p < LP; ++p) from_window(rq[p]);
    } else {
#pragma unroll
      for (int p = 0; p < LP; ++p) {
        const int l = p / 4;
        const float4 r = rq[p];
        if (wa.base[l] >= 0) {  // block-uniform
          from_window(r);
        } else {
          const int code = __float_as_int(r.x);
          const float lh = r.y, lw = r.z, a = r.w;
          const float a_hh = a * (1.f - lh), a_lh = a * lh, hw = 1.f - lw;
          const VT* p0 = value_j + (ptrdiff_t)(code & ~31);
          const VT* p2 = p0 + (ptrdiff_t)(lv.W[l] * M32);
          float v[C];
          if (code & 1) {
            RT::load(p0, v);
            const float w = a_hh * hw;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(w, v[c], acc[c]);
          }
          if (code & 2) {
            RT::load(p0 + M32, v);
            const float w = a_hh * lw;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(w, v[c], acc[c]);
          }
          if (code & 4) {
            RT::load(p2, v);
            const float w = a_lh * hw;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(w, v[c], acc[c]);
          }
          if (code & 8) {
            RT::load(p2 + M32, v);
            const float w = a_lh * lw;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(w, v[c], acc[c]);
          }
        }
      }
    }
    RT::store_stream(out + qm * 32 + j * C, acc);
  }
#ifdef MSDA_WIN_TIMING
  __syncthreads();
  if (threadIdx.x == 0 && !all_win) { atomicAdd(&g_win_timing[4], (unsigned long long)(clock64() - tphase)); atomicAdd(&g_win_timing[5], 1ull); }
  WIN_T(3, tphase);  // gather
  if (threadIdx.x == 0) atomicAdd(&g_win_timing[7], 1ull);
#endif
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// How the 4 lanes of a sorted-pass group cover the 32 channels of a row: lane sj owns two float4 chunks
// (chunk = 4 channels) of every fp32 row (grad_out, grad_value) and the same channels of the value rows.
//   fp32 value: chunks sj and sj + 4; odd groups take them in the opposite order so that the two groups
//               of a quarter-warp hit different bank halves (conflict-free LDS.128).
//   bf16 value: one 16-byte read = channels 8 sj .. 8 sj + 7 = chunks 2 sj, 2 sj + 1.
template <typename VT>
struct SortLane;
template <>
struct SortLane<float> {
  static __device__ __forceinline__ int chunk_a(int sg, int sj) { return sj + ((sg & 1) ? 4 : 0); }
  static __device__ __forceinline__ int chunk_b(int sg, int sj) { return sj + ((sg & 1) ? 0 : 4); }
  static __device__ __forceinline__ int pool_offset(int sg, int sj) { return chunk_a(sg, sj) * 16; }
  static __device__ __forceinline__ void lds(const unsigned char* p, float (&v)[8]) {
    const float4 lo = *reinterpret_cast<const float4*>(p);
    const float4 hi = *reinterpret_cast<const float4*>(reinterpret_cast<uintptr_t>(p) ^ 64);  // the other half of the row
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  }
};
template <>
struct SortLane<__nv_bfloat16> {
  static __device__ __forceinline__ int chunk_a(int, int sj) { return 2 * sj; }
  static __device__ __forceinline__ int chunk_b(int, int sj) { return 2 * sj + 1; }
  static __device__ __forceinline__ int pool_offset(int, int sj) { return sj * 16; }
  static __device__ __forceinline__ void lds(const unsigned char* p, float (&v)[8]) {
    RowTraits<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(p), v);
  }
};

// Reduction of `C` floats held by each of the G lanes of a group into grad_value row `off`.
template <int C>
__device__ __forceinline__ void win_red_row(float* gvalue_j, const int off, const float (&a)[C]) {
  if (off >= 0) {
#pragma unroll
    for (int c = 0; c < C; c += 4) red_add_f4(gvalue_j + off + c, a[c], a[c + 1], a[c + 2], a[c + 3]);
  }
}

#ifndef MSDA_WIN_BWD_MINBLOCKS
#define MSDA_WIN_BWD_MINBLOCKS 2
#endif

// Shared-memory working set of one backward tile (one "buffer set").
template <class Cfg>
struct WinBwdSmem {
  unsigned char* pool;
  float4* rec;
  int* bb;
  float* go_s;
  int* hist;               // per-cell counts, then exclusive offsets
  int* rowoff;
  unsigned short* sorted;  // sample ids sorted by cell
  int* misc;               // [0,16) warp totals [16] total [20,28) window base row per level (-1: direct)
  float* lvf;              // [0,8) (float)W_l  [8,16) (float)H_l
  unsigned char* inflag;   // per (level slot, query): bit i = point i passed the range test
  unsigned short* wcnt;    // deterministic mode only: [warp][cell] counts, then exclusive bases over the warps
  float2* stats;           // fused prologue only: softmax (max, sum) of every query of the tile
  __device__ __forceinline__ explicit WinBwdSmem(unsigned char* base)
      : pool(base),
        rec(reinterpret_cast<float4*>(base + Cfg::POOL_BYTES)),
        bb(reinterpret_cast<int*>(base + Cfg::POOL_BYTES + Cfg::REC_BYTES)),
        go_s(reinterpret_cast<float*>(base + Cfg::OFF_GO)),
        hist(reinterpret_cast<int*>(base + Cfg::OFF_HIST)),
        rowoff(reinterpret_cast<int*>(base + Cfg::OFF_ROWOFF)),
        sorted(reinterpret_cast<unsigned short*>(base + Cfg::OFF_SORTED)),
        misc(reinterpret_cast<int*>(base + Cfg::OFF_SORTED + Cfg::SORTED_BYTES)),
        lvf(reinterpret_cast<float*>(base + Cfg::OFF_SORTED + Cfg::SORTED_BYTES) + 32),
        inflag(base + Cfg::OFF_SORTED + Cfg::SORTED_BYTES + 192),
        stats(reinterpret_cast<float2*>(base + Cfg::OFF_SORTED + Cfg::SORTED_BYTES + 192 + Cfg::INFLAG_BYTES)),
        wcnt(reinterpret_cast<unsigned short*>(base + ((Cfg::BWD_SMEM + 15) / 16) * 16)) {}
};

struct WinBwdArgs {
  const void* grad_out;
  const void* value;
  const float* loc;
  const float* attw;
  float* grad_value;
  float* grad_loc;
  float* grad_attw;
  const int* order;
  int order_len, S, M, Lq;
  // deterministic mode only: fixed-point accumulators (one per grad_value element) and the bit patterns of
  // max|grad_out|, max|attn_weight| that fix their scale
  long long* gv64;
  const unsigned* maxbits;
  // fused prologue (ref_dim != 0): loc / attw are the raw offsets / logits, grad_loc / grad_attw receive the
  // gradients of the raw tensors
  MsdaFused fz;
};

// Deterministic mode: scale (a power of two) that maps any sum of up to Lq*L*P products weight * grad_out,
// |weight| <= max|attn_weight|, into 62 bits.  Integer addition is associative, so the order in which blocks
// add their (canonically computed) partial sums no longer matters.
__device__ __forceinline__ int win_det_shift(const unsigned* maxbits, const int Lq, const int LP) {
  const int e_go = (int)((maxbits[0] >> 23) & 0xff) - 126;  // max|grad_out| < 2^e_go
  const int e_a = (int)((maxbits[1] >> 23) & 0xff) - 126;
  const int e_n = 32 - __clz(max(Lq * LP, 1));                // Lq*L*P < 2^e_n
  return max(-120, min(120, 62 - (e_go + e_a + e_n)));        // the scale 2^shift is an fp32 normal
}
__device__ __forceinline__ float win_det_scale(const unsigned* maxbits, const int Lq, const int LP) {
  return __uint_as_float((unsigned)(127 + win_det_shift(maxbits, Lq, LP)) << 23);
}
// v * 2^shift is exact in fp32 (no overflow by construction), so one FMUL and one F2I do the conversion
__device__ __forceinline__ void win_det_add(long long* p, const float v, const float scale) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * scale));
}
// Position of channel c inside a row of the fixed-point accumulator array (a private workspace, so its
// layout is ours to choose): channel 4*chunk + i -> 8*i + chunk.  Lanes that hold the same element i of
// consecutive chunks — which is how both the sorted pass and the direct pass distribute a row — then add
// into consecutive accumulators, i.e. one 32-byte sector per four lanes (L2 retires atomics per sector).
__device__ __forceinline__ int win_det_pos(const int c) { return 8 * (c & 3) + (c >> 2); }
// 4x4 transpose over the 4 lanes of a quad (q = lane & 3): afterwards a[L] is what lane L of the quad held in
// a[q].  Lets the four lanes of one 64-bit reduction instruction hit four CONSECUTIVE accumulators (one 32-byte
// sector) instead of four sectors — L2 retires atomics per sector.
__device__ __forceinline__ void quad_transpose4(float (&a)[4], const int q, const unsigned mask) {
  {
    const bool hi = q & 1;
    const float r0 = __shfl_xor_sync(mask, hi ? a[0] : a[1], 1), r1 = __shfl_xor_sync(mask, hi ? a[2] : a[3], 1);
    if (hi) { a[0] = r0; a[2] = r1; } else { a[1] = r0; a[3] = r1; }
  }
  {
    const bool hi = q & 2;
    const float r0 = __shfl_xor_sync(mask, hi ? a[0] : a[2], 2), r1 = __shfl_xor_sync(mask, hi ? a[1] : a[3], 2);
    if (hi) { a[0] = r0; a[1] = r1; } else { a[2] = r0; a[3] = r1; }
  }
}

// Producer half of a backward tile: decode, windows (staged with cp.async), records, counting sort,
// grad_out rows.  `t` = thread index inside the kWinThreads-wide group that runs it, `sync` its barrier.
// On return everything the consumer half needs is in the buffer set and visible to the group.
#ifndef MSDA_WIN_MATCH_RANK
#define MSDA_WIN_MATCH_RANK 0  // 1: match-based (scheduling-independent) sort ranks in the atomic mode too
#endif
template <typename VT, int kL, int kWinPool, bool kDet, bool kFused, class Sync>
__device__ __forceinline__ void win_bwd_produce(const WinBwdSmem<WinCfg<VT, kL, kWinPool>>& sm, const WinBwdArgs& ar,
                                                const MsdaLevels& lv, const int tile, const int m, const int b,
                                                const int t, const Sync sync) {
  using Cfg = WinCfg<VT, kL, kWinPool>;
  using RT = RowTraits<VT>;
  constexpr int LP = Cfg::LP, G = RT::G, C = RT::C;
  constexpr bool kRank = kDet || (MSDA_WIN_MATCH_RANK != 0);
  const int M = ar.M, Lq = ar.Lq;
  const int warp = t >> 5, lane = t & 31;
  const VT* grad_out = static_cast<const VT*>(ar.grad_out);
  const VT* value_img = static_cast<const VT*>(ar.value) + (size_t)b * ar.S * M * 32;
  const int* order = ar.order;
  const int order_len = ar.order_len;
  long long tphase = 0;

  const int dql = t & (kWinTileQ - 1), dslot = t / kWinTileQ;
  int dq = -1;
  {
    const int oslot = tile * kWinTileQ + dql;
    if (oslot < order_len) dq = order ? order[oslot] : oslot;
  }
  const size_t dqm = ((size_t)b * Lq + (dq >= 0 ? dq : 0)) * M + m;
  int qpf = -1;
  {
    const int pslot = (tile + kWinPrefetchTiles) * kWinTileQ + dql;
    if (pslot < order_len) qpf = order ? order[pslot] : pslot;
  }
  const size_t qm_pf = ((size_t)b * Lq + (qpf >= 0 ? qpf : 0)) * M + m;
  if (qpf >= 0 && dslot == 0) prefetch_l2(grad_out + qm_pf * 32);
  // grad_out rows of the tile: loads issued now, parked in shared memory (as fp32) after the front end, so
  // that their latency overlaps the decode's own loads instead of preceding them
  constexpr int GO_ITERS = kWinTileQ * G / kWinThreads;
  static_assert(kWinTileQ * G % kWinThreads == 0, "grad_out staging covers the tile in whole iterations");
  float gvreg[GO_ITERS][C];
#pragma unroll
  for (int it = 0; it < GO_ITERS; ++it) {
    const int i = t + it * kWinThreads;
    const int gql = i / G, gj = i % G;
    const int oslot = tile * kWinTileQ + gql;
    int gq = -1;
    if (oslot < order_len) gq = order ? order[oslot] : oslot;
#pragma unroll
    for (int c = 0; c < C; ++c) gvreg[it][c] = 0.f;
    if (gq >= 0) RT::load_stream(grad_out + (((size_t)b * Lq + gq) * M + m) * 32 + gj * C, gvreg[it]);
  }
  WinAlloc<kL> wa;
  WinPoint pts[Cfg::NLV][4];
  int rank[Cfg::NLV][4];
  win_front_end<VT, kL, kWinPool, true, kRank>(t, sync, sm.wcnt, value_img, ar.loc, ar.attw, dq, dqm,
                                              kFused ? ar.fz : MsdaFused{nullptr, 0},
                                              (size_t)b * Lq + (dq >= 0 ? dq : 0), kFused ? sm.stats + dql : nullptr, qpf, qm_pf, m, M, lv, sm.pool, sm.rec,
                                        sm.bb, sm.rowoff, sm.hist, wa, pts, rank, tphase);
#pragma unroll
  for (int it = 0; it < GO_ITERS; ++it) {
    const int i = t + it * kWinThreads;
#pragma unroll
    for (int c = 0; c < C; c += 4)
      *reinterpret_cast<float4*>(sm.go_s + (i / G) * 32 + (i % G) * C + c) =
          make_float4(gvreg[it][c], gvreg[it][c + 1], gvreg[it][c + 2], gvreg[it][c + 3]);
  }
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li)
    sm.inflag[(li * 4 + dslot) * kWinTileQ + dql] =
        (unsigned char)((pts[li][0].in ? 1 : 0) | (pts[li][1].in ? 2 : 0) | (pts[li][2].in ? 4 : 0) | (pts[li][3].in ? 8 : 0));
  if (t < 2) sm.rowoff[kWinPool + t] = -1;
  if (t >= 32 && t < 32 + kL) {
    sm.lvf[t - 32] = (float)lv.W[t - 32];
    sm.lvf[8 + t - 32] = (float)lv.H[t - 32];
#pragma unroll
    for (int l = 0; l < kL; ++l)
      if (l == t - 32) sm.misc[20 + l] = wa.base[l];
  }
  sync();  // hist complete, records visible
  if (kRank) {
    // per cell: the warps' counts -> exclusive bases over the warps, their sum -> hist
    for (int c = t; c < Cfg::HIST_N; c += kWinThreads) {
      int run = 0;
#pragma unroll
      for (int w = 0; w < kWinThreads / 32; ++w) {
        const int n = sm.wcnt[w * Cfg::HIST_N + c];
        sm.wcnt[w * Cfg::HIST_N + c] = (unsigned short)run;
        run += n;
      }
      sm.hist[c] = run;
    }
    sync();
  }

  // ---- exclusive scan of the per-cell counts, in place -----------------------------------------
  {
    int v[Cfg::SPT], sum = 0;
#pragma unroll
    for (int k = 0; k < Cfg::SPT; ++k) { v[k] = sm.hist[t * Cfg::SPT + k]; sum += v[k]; }
    int inc = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= s) inc += n;
    }
    if (lane == 31) sm.misc[warp] = inc;
    sync();
    int run = inc - sum;
    for (int w = 0; w < warp; ++w) run += sm.misc[w];
#pragma unroll
    for (int k = 0; k < Cfg::SPT; ++k) { sm.hist[t * Cfg::SPT + k] = run; run += v[k]; }
    if (t == kWinThreads - 1) sm.misc[16] = run;
    sync();
  }
  // ---- place the sample ids at their sorted positions ---------------------------------------------
#pragma unroll
  for (int li = 0; li < Cfg::NLV; ++li) {
    const int l = dslot + 4 * li;
    if (l < kL) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (rank[li][i] >= 0) {
          const int code = __float_as_int(sm.rec[dql * Cfg::REC_STRIDE + l * 4 + i].x);
          WIN_CHECK(sm.hist[code & 0xffff] + rank[li][i] >= 0 && sm.hist[code & 0xffff] + rank[li][i] < sm.misc[16]);
          const int wbase = kRank ? (int)sm.wcnt[warp * Cfg::HIST_N + (code & 0xffff)] : 0;
          sm.sorted[sm.hist[code & 0xffff] + wbase + rank[li][i]] = (unsigned short)(dql * LP + l * 4 + i);
        }
    }
  }
  cp_async_wait_all();
  sync();  // windows, sorted list, grad_out rows are in shared memory
}

// Consumer half: sorted pass, direct pass, write-out (see the file header).
template <typename VT, int kL, int kWinPool, bool kDet, bool kFused, class Sync>
__device__ __forceinline__ void win_bwd_consume(const WinBwdSmem<WinCfg<VT, kL, kWinPool>>& sm, const WinBwdArgs& ar,
                                                const MsdaLevels& lv, const int tile, const int m, const int b,
                                                const int t, const Sync sync) {
  using Cfg = WinCfg<VT, kL, kWinPool>;
  using RT = RowTraits<VT>;
  constexpr int LP = Cfg::LP, G = RT::G, C = RT::C, GPW = 32 / G, ROWB = Cfg::ROWB;
  constexpr int NG = (kWinThreads / 32) * GPW;  // lane groups per block
  const int M = ar.M, Lq = ar.Lq, S = ar.S;
  const int M32 = M * 32;
  const int warp = t >> 5, lane = t & 31;
  const size_t img = (size_t)b * S * M32;
  const VT* value_img = static_cast<const VT*>(ar.value) + img;
  float* grad_value = ar.grad_value;
  unsigned char* pool = sm.pool;
  float4* rec = sm.rec;
  float* go_s = sm.go_s;
  int* rowoff = sm.rowoff;
  unsigned short* sorted = sm.sorted;
  int* misc = sm.misc;
  float* lvf = sm.lvf;
  int lbase[kL];
#pragma unroll
  for (int l = 0; l < kL; ++l) lbase[l] = misc[20 + l];
  const float dscale = kDet ? win_det_scale(ar.maxbits, Lq, LP) : 0.f;
  long long* gv64 = kDet ? ar.gv64 + img : nullptr;

  // ---- sorted pass: one 4-lane group (8 channels per lane) per contiguous chunk of the cell-sorted samples ----
  {
    using SL = SortLane<VT>;
    constexpr int SG = 4, SC = 8, SNG = kWinThreads / SG;
    const int sg = lane >> 2, sj = lane & 3;
    const int oA = SL::chunk_a(sg, sj) * 16, oB = SL::chunk_b(sg, sj) * 16;  // byte offsets in a 128-byte fp32 row
    const unsigned char* pool_v = pool + SL::pool_offset(sg, sj);
    float* gvalue_a = grad_value + img + oA / 4;
    float* gvalue_b = grad_value + img + oB / 4;
    const unsigned char* go_a = reinterpret_cast<const unsigned char*>(go_s) + oA;
    const unsigned char* go_b = reinterpret_cast<const unsigned char*>(go_s) + oB;
    const int total = misc[16];
    const int chunk = (((total + SNG - 1) / SNG) + 3) & ~3;  // multiple of the batch size
    const int gi = warp * 8 + sg;
    const int i0 = min(total, gi * chunk), i1 = min(total, i0 + chunk);
    const unsigned gmask = 0xfu << (sg * 4);
    // channel pairs: every FMA below is one FFMA2
    constexpr int SP = SC / 2;
    float2 V00[SP], V01[SP], V10[SP], V11[SP], A0[SP], A1[SP], B0[SP], B1[SP];
    const float2 zero2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < SP; ++c) V00[c] = V01[c] = V10[c] = V11[c] = A0[c] = A1[c] = B0[c] = B1[c] = zero2;
    int cur0 = -2, cur1 = -2;
    auto flush = [&](const int row, const float2 (&acc)[SP]) {
      WIN_CHECK(row >= 0 && row < kWinPool + 2);
      const int off = rowoff[row];
      WIN_CHECK(off < 0 || (off % 32 == 0 && off / 32 < S * M));
      if (off >= 0) {
        if (kDet) {
          // accumulator layout (win_det_pos): channel 4*chunk + i sits at 8*i + chunk, so the four lanes of
          // one reduction instruction (same i, chunks sj .. sj+3) hit one 32-byte sector
          long long* pa = gv64 + off + SL::chunk_a(sg, sj);
          long long* pb = gv64 + off + SL::chunk_b(sg, sj);
          win_det_add(pa, acc[0].x, dscale); win_det_add(pa + 8, acc[0].y, dscale);
          win_det_add(pa + 16, acc[1].x, dscale); win_det_add(pa + 24, acc[1].y, dscale);
          win_det_add(pb, acc[2].x, dscale); win_det_add(pb + 8, acc[2].y, dscale);
          win_det_add(pb + 16, acc[3].x, dscale); win_det_add(pb + 24, acc[3].y, dscale);
        } else {
          red_add_f4(gvalue_a + off, acc[0].x, acc[0].y, acc[1].x, acc[1].y);
          red_add_f4(gvalue_b + off, acc[2].x, acc[2].y, acc[3].x, acc[3].y);
        }
      }
    };
    auto load_row = [&](const int row, float2 (&v)[SP]) {
      float t8[SC];
      SL::lds(pool_v + row * ROWB, t8);
#pragma unroll
      for (int c = 0; c < SP; ++c) v[c] = make_float2(t8[2 * c], t8[2 * c + 1]);
    };
    auto rec_slot = [&](const int sid) { const int sq = sid / LP; return sq * Cfg::REC_STRIDE + (sid - sq * LP); };
    float4 rnext = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 < i1) rnext = rec[rec_slot(sorted[i0])];
    for (int ib = i0; ib < i1; ib += 4) {
      const uint2 packed = *reinterpret_cast<const uint2*>(sorted + ib);  // 4 sample ids (ib is a multiple of 4)
      const int nb = min(4, i1 - ib);
      int nsid = 0;
      if (ib + 4 < i1) nsid = sorted[ib + 4];
      float pgx[4], pgy[4], pga[4], aws[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pgx[u] = pgy[u] = pga[u] = aws[u] = 0.f;
        if (u < nb) {  // group-uniform
          const int sid = (int)(((u < 2 ? packed.x : packed.y) >> ((u & 1) * 16)) & 0xffffu);
          const int sq = sid / LP, l = (sid - sq * LP) >> 2;
          const float4 r = rnext;
          // grad_out row of the sample's query; the next sample's record is fetched one step ahead
          const float4 ga4 = *reinterpret_cast<const float4*>(go_a + sq * 128);
          const float4 gb4 = *reinterpret_cast<const float4*>(go_b + sq * 128);
          {
            const int sidn = u == 3 ? nsid
                                    : (int)((((u + 1) < 2 ? packed.x : packed.y) >> (((u + 1) & 1) * 16)) & 0xffffu);
            if (u + 1 < nb || (u == 3 && ib + 4 < i1)) rnext = rec[rec_slot(sidn)];
          }
          const int code = __float_as_int(r.x);
          const int row0 = code & 0xffff, row1 = code >> 16;
          WIN_CHECK(sid < kWinTileQ * LP && row0 >= cur0 && row1 + 1 < kWinPool && row1 > row0);
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
            load_row(row0, V00);
            load_row(row0 + 1, V01);
            load_row(row1, V10);
            load_row(row1 + 1, V11);
            cur0 = row0; cur1 = row1;
          }
          const float lh = r.y, lw = r.z, a = r.w;
          if (kFused) aws[u] = a;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float2 go[SP] = {make_float2(ga4.x, ga4.y), make_float2(ga4.z, ga4.w), make_float2(gb4.x, gb4.y),
                                 make_float2(gb4.z, gb4.w)};
          // grad_value partial sums (cuh:125,134,143,152): independent of the value rows, so they cover
          // the latency of a window reload
          const float c00 = hh * hw, c01 = hh * lw, c10 = lh * hw, c11 = lh * lw;
          const float w00 = c00 * a, w01 = c01 * a, w10 = c10 * a, w11 = c11 * a;
          const float2 w00p = make_float2(w00, w00), w01p = make_float2(w01, w01), w10p = make_float2(w10, w10),
                       w11p = make_float2(w11, w11);
#pragma unroll
          for (int c = 0; c < SP; ++c) {
            A0[c] = ffma2(w00p, go[c], A0[c]); A1[c] = ffma2(w01p, go[c], A1[c]);
            B0[c] = ffma2(w10p, go[c], B0[c]); B1[c] = ffma2(w11p, go[c], B1[c]);
          }
          float2 e00 = zero2, e01 = zero2, e10 = zero2, e11 = zero2;
#pragma unroll
          for (int c = 0; c < SP; ++c) {
            e00 = ffma2(go[c], V00[c], e00); e01 = ffma2(go[c], V01[c], e01);
            e10 = ffma2(go[c], V10[c], e10); e11 = ffma2(go[c], V11[c], e11);
          }
          const float d00 = e00.x + e00.y, d01 = e01.x + e01.y, d10 = e10.x + e10.y, d11 = e11.x + e11.y;
          // grad_attn_weight (cuh:156), grad_sampling_loc (cuh:157-158): this lane's channels
          pga[u] = fmaf(c00, d00, fmaf(c01, d01, fmaf(c10, d10, c11 * d11)));
          pgx[u] = (a * lvf[l]) * fmaf(hh, d01 - d00, lh * (d11 - d10));
          pgy[u] = (a * lvf[8 + l]) * fmaf(hw, d10 - d00, lw * (d11 - d01));
        }
      }
      const float gx = group_reduce_scatter<4>(pgx, sj, gmask);
      const float gy = group_reduce_scatter<4>(pgy, sj, gmask);
      const float ga = group_reduce_scatter<4>(pga, sj, gmask);
      if (sj < nb) {  // lane sj owns sample ib + sj: park its gradients in the sample's record slot
        const int sid = (int)(((sj < 2 ? packed.x : packed.y) >> ((sj & 1) * 16)) & 0xffffu);
        // .w keeps the sample's weight (the fused prologue's softmax backward needs it)
        rec[rec_slot(sid)] = make_float4(gx, gy, ga, !kFused ? 0.f : sj == 0 ? aws[0] : sj == 1 ? aws[1] : sj == 2 ? aws[2] : aws[3]);
      }
    }
    if (cur0 >= 0) {
      flush(cur0, A0);
      flush(cur1, B0);
      flush(cur0 + 1, A1);
      flush(cur1 + 1, B1);
    }
  }
  const int g = lane / G, j = lane % G;
  float* gvalue_j = grad_value + img + j * C;

  // ---- direct pass: levels that did not get a window, query-major from global memory -------------------
  bool all_win = true;
#pragma unroll
  for (int l = 0; l < kL; ++l) all_win = all_win && lbase[l] >= 0;
  if (!all_win) {
    const VT* value_j = value_img + j * C;
    for (int ql = warp * GPW + g; ql < kWinTileQ; ql += NG) {
      float go[C];
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 gq = *reinterpret_cast<const float4*>(go_s + ql * 32 + j * C + c);
        go[c] = gq.x; go[c + 1] = gq.y; go[c + 2] = gq.z; go[c + 3] = gq.w;
      }
#pragma unroll
      for (int l = 0; l < kL; ++l) {
        if (lbase[l] >= 0) continue;  // block-uniform
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4* slot = rec + ql * Cfg::REC_STRIDE + l * 4 + i;
          const float4 r = *slot;
          const int code = __float_as_int(r.x);
          const float lh = r.y, lw = r.z, a = r.w;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float a_hh = a * hh, a_lh = a * lh;
          const ptrdiff_t o0 = (ptrdiff_t)(code & ~31);
          const ptrdiff_t o2 = o0 + (ptrdiff_t)(lv.W[l] * M32);
          float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (code & (1 << k)) {
              const ptrdiff_t o = ((k & 2) ? o2 : o0) + ((k & 1) ? M32 : 0);
              float v[C];
              RT::load(value_j + o, v);
              const float tt = ((k & 2) ? a_lh : a_hh) * ((k & 1) ? lw : hw);
              if (kDet) {
#pragma unroll
                for (int c = 0; c < C; ++c)  // lane j holds channels j*C + c: chunk (j*C + c) / 4, element (j*C + c) % 4
                  win_det_add(gv64 + o + win_det_pos(j * C + c), tt * go[c], dscale);
              } else {
#pragma unroll
                for (int c = 0; c < C; c += 4)
                  red_add_f4(gvalue_j + o + c, tt * go[c], tt * go[c + 1], tt * go[c + 2], tt * go[c + 3]);
              }
              float sdot = 0.f;
#pragma unroll
              for (int c = 0; c < C; ++c) sdot = fmaf(go[c], v[c], sdot);
              d[k] = sdot;
            }
          }
          float ga = hh * (hw * d[0] + lw * d[1]) + lh * (hw * d[2] + lw * d[3]);
          float gx = (a * (float)lv.W[l]) * (hh * (d[1] - d[0]) + lh * (d[3] - d[2]));
          float gy = (a * (float)lv.H[l]) * (hw * (d[2] - d[0]) + lw * (d[3] - d[1]));
#pragma unroll
          for (int s = G / 2; s >= 1; s >>= 1) {
            ga += __shfl_xor_sync(0xffffffffu, ga, s);
            gx += __shfl_xor_sync(0xffffffffu, gx, s);
            gy += __shfl_xor_sync(0xffffffffu, gy, s);
          }
          __syncwarp();
          if (j == 0) *slot = make_float4(gx, gy, ga, a);
        }
      }
    }
  }
  sync();  // every sample's gradients are parked in its record slot

  // ---- write-out: thread <-> (level slot, query) as in the decode ------------------------------------
  const int dql = t & (kWinTileQ - 1), dslot = t / kWinTileQ;
  int dq = -1;
  {
    const int oslot = tile * kWinTileQ + dql;
    if (oslot < ar.order_len) dq = ar.order ? ar.order[oslot] : oslot;
  }
  if (dq >= 0) {
    const size_t dqm = ((size_t)b * Lq + dq) * M + m;
#pragma unroll
    for (int li = 0; li < Cfg::NLV; ++li) {
      const int l = dslot + 4 * li;
      if (l < kL) {
        const int fl = sm.inflag[(li * 4 + dslot) * kWinTileQ + dql];
        float4 r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r[i] = rec[dql * Cfg::REC_STRIDE + l * 4 + i];
          if (!(fl & (1 << i))) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // skipped sample: slot still holds its record
        }
        if (kFused) {
          // fused prologue: gradients of the raw offsets and logits.  softmax backward needs
          // dot = sum_j a_j * dL/da_j over all L*P samples of (query, head): the other levels' dL/da sit in
          // their record slots, the weights are recomputed from the logits and the saved (max, sum)
          float dot = 0.f;
          for (int sl = 0; sl < Cfg::NLV * 4; ++sl) {
            const int l2 = (sl & 3) + 4 * (sl >> 2);
            if (l2 < kL) {
              const int f2 = sm.inflag[sl * kWinTileQ + dql];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 o = rec[dql * Cfg::REC_STRIDE + l2 * 4 + i];  // (gx, gy, dL/da, a) of a processed sample
                if (f2 & (1 << i)) dot = fmaf(o.w, o.z, dot);
              }
            }
          }
          const float2 st = sm.stats[dql];
          const float4 x = ld_stream_f4(ar.attw + dqm * LP + l * 4);
          const float xs[4] = {x.x, x.y, x.z, x.w};
          const float* rp = ar.fz.ref + (((size_t)b * Lq + dq) * kL + l) * ar.fz.ref_dim;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // a skipped sample has no result slot: its weight is recomputed (its dL/da is 0, not its dL/dlogit)
            const float a = (fl & (1 << i)) ? rec[dql * Cfg::REC_STRIDE + l * 4 + i].w : __expf(xs[i] - st.x) / st.y;
            // d loc / d offset, in autograd's operation order: g / W  |  ((g * 0.5) * wh) / P
            const float gx = ar.fz.ref_dim == 2 ? r[i].x / (float)lv.W[l] : r[i].x * 0.5f * rp[2] / 4.f;
            const float gy = ar.fz.ref_dim == 2 ? r[i].y / (float)lv.H[l] : r[i].y * 0.5f * rp[3] / 4.f;
            r[i] = make_float4(gx, gy, a * (r[i].z - dot), 0.f);
          }
        }
        float* gl = ar.grad_loc + (dqm * LP + l * 4) * 2;
        st_stream_f4(gl, make_float4(r[0].x, r[0].y, r[1].x, r[1].y));
        st_stream_f4(gl + 4, make_float4(r[2].x, r[2].y, r[3].x, r[3].y));
        st_stream_f4(ar.grad_attw + dqm * LP + l * 4, make_float4(r[0].z, r[1].z, r[2].z, r[3].z));
      }
    }
  }
}

// One block = one tile x one head: produce, then consume.  kDet: deterministic grad_value (canonical order
// inside the block, order-independent fixed-point accumulation across blocks; see msda_capi.cu).
template <typename VT, int kL, int kM, bool kDet, bool kFused>
__global__ void __launch_bounds__(kWinThreads, MSDA_WIN_BWD_MINBLOCKS)
msda_bwd_d32_win_kernel(const WinBwdArgs ar, const __grid_constant__ MsdaLevels lv) {
  constexpr int kWinPool = kWinPoolBwd;
  using Cfg = WinCfg<VT, kL, kWinPool>;
  static_assert(kWinTileQ * 4 == kWinThreads, "decode maps 4 threads to a query");
  extern __shared__ __align__(128) unsigned char smraw[];
  const WinBwdSmem<Cfg> sm(smraw);
  WinBwdArgs a = ar;
  if (kM) a.M = kM;
  const int m = blockIdx.x % a.M, tile = blockIdx.x / a.M, b = blockIdx.y;
  win_bwd_produce<VT, kL, kWinPool, kDet, kFused>(sm, a, lv, tile, m, b, (int)threadIdx.x, BlockSync{});
  // the front end writes nothing to global memory; everything from here on may reduce into grad_value, which the
  // preceding kernel on the stream zero-fills when this kernel was allowed to start early (programmatic dependent
  // launch, msda_capi.cu); otherwise the wait returns at once
  if (!kDet) asm volatile("griddepcontrol.wait;" ::: "memory");
  win_bwd_consume<VT, kL, kWinPool, kDet, kFused>(sm, a, lv, tile, m, b, (int)threadIdx.x, BlockSync{});
}

// Deterministic mode helpers: max|x| of a tensor as a float bit pattern (non-negative floats order like
// unsigned ints; a NaN / Inf input ends up >= 0x7f800000), and the final fixed-point -> fp32 conversion.
template <typename T>
__global__ void __launch_bounds__(256)
msda_maxabs_kernel(const T* __restrict__ go, const size_t n_go, const float* __restrict__ aw, const size_t n_aw, unsigned* out) {
  // 16-byte loads; both tensors have a multiple of 16 bytes (32 channels / 4 points per row)
  constexpr int EPV = 16 / (int)sizeof(T);
  const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nth = (size_t)gridDim.x * 256;
  unsigned m_go = 0, m_aw = 0;
  for (size_t i = tid; i < n_go / EPV; i += nth) {
    float v[EPV];
    RowTraits<T>::load_stream(go + i * EPV, v);
#pragma unroll
    for (int c = 0; c < EPV; ++c) m_go = max(m_go, __float_as_uint(fabsf(v[c])));
  }
  for (size_t i = tid; i < n_aw / 4; i += nth) {
    const float4 v = ld_stream_f4(aw + i * 4);
    m_aw = max(max(m_aw, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
  }
  m_go = __reduce_max_sync(0xffffffffu, m_go);
  m_aw = __reduce_max_sync(0xffffffffu, m_aw);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, m_go);
    atomicMax(out + 1, m_aw);
  }
}
__global__ void __launch_bounds__(256) msda_fixed_to_float_kernel(const long long* __restrict__ acc, float* __restrict__ out,
                                                                  const size_t n, const unsigned* __restrict__ maxbits,
                                                                  const int Lq, const int LP) {
  const bool finite = maxbits[0] < 0x7f800000u && maxbits[1] < 0x7f800000u;
  const double inv = exp2(-(double)win_det_shift(maxbits, Lq, LP));
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    out[i] = finite ? (float)((double)acc[(i & ~(size_t)31) + win_det_pos((int)(i & 31))] * inv) : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------------------
// backward, persistent and warp-specialised: one block per SM with two kWinThreads-wide groups.  The
// producer group runs the front end of tile i+1 into one buffer set while the consumer group runs the
// sorted pass of tile i out of the other, so the front end's load latencies and barriers overlap the
// sorted pass's arithmetic instead of alternating with it.  Hand-off through two pairs of mbarriers
// (full / empty per buffer set), barriers inside a group are named barriers (bar.sync id, 256).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, const int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, const unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned done = 0;
#ifdef MSDA_WIN_CHECKS
  long long spins = 0;
#endif
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#ifdef MSDA_WIN_CHECKS
    if (++spins > (1ll << 26)) asm volatile("trap;");  // a lost hand-off must not hang the GPU
#endif
  }
}

template <typename VT, int kL, int kM>
__global__ void __launch_bounds__(2 * kWinThreads, 1)
msda_bwd_d32_ws_kernel(const WinBwdArgs ar, const __grid_constant__ MsdaLevels lv, const int tiles, const int batch) {
  constexpr int kWinPool = kWinPoolBwd;
  using Cfg = WinCfg<VT, kL, kWinPool>;
  extern __shared__ __align__(128) unsigned char smraw[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smraw + 2 * Cfg::BWD_SET_BYTES);  // full[2], empty[2]
  WinBwdArgs a = ar;
  if (kM) a.M = kM;
  const int role = threadIdx.x / kWinThreads, t = threadIdx.x % kWinThreads;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int total = tiles * a.M * batch;
  int it = 0;
  for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
    const int set = it & 1;
    const unsigned use = (unsigned)(it >> 1);  // how many times this buffer set has been used before
    const WinBwdSmem<Cfg> sm(smraw + set * Cfg::BWD_SET_BYTES);
    const int m = w % a.M, tb = w / a.M;
    const int tile = tb % tiles, b = tb / tiles;
    long long tph = clock64();
    (void)tph;
    if (role == 0) {
      if (use > 0) mbar_wait(bars + 2 + set, (use - 1) & 1);  // the consumer released the set
      if (t == 0) WIN_T(11, tph);  // producer waiting for a free buffer set
      win_bwd_produce<VT, kL, kWinPool, false, false>(sm, a, lv, tile, m, b, t, GroupSync<1>{});
      if (t == 0) WIN_T(8, tph);   // produce
      if (t == 0) mbar_arrive(bars + set);
    } else {
      mbar_wait(bars + set, use & 1);
      if (t == 0) WIN_T(9, tph);   // consumer waiting for a full buffer set
      win_bwd_consume<VT, kL, kWinPool, false, false>(sm, a, lv, tile, m, b, t, GroupSync<2>{});
      GroupSync<2>{}();  // every consumer thread is done with the set
      if (t == 0) WIN_T(10, tph);  // consume
#ifdef MSDA_WIN_TIMING
      if (t == 0) atomicAdd(&g_win_timing[15], 1ull);
#endif
      if (t == 0) mbar_arrive(bars + 2 + set);
    }
  }
}

}  // namespace msda
