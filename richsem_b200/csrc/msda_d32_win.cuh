// msda_d32_win.cuh — "window" backward for head_dim 32 (encoder-sized problems): the value rows a tile of
// queries samples are staged ONCE in shared memory, and the tile's samples are walked sorted by window cell
// so that value rows AND grad_value partial sums live in registers.
//
// Why (measured on B200, scratch/gather_bw.cu, scratch/smem_atomics.cu, scratch/tma_red.cu):
//   * L2 retires scattered fp32 row reductions at 5.8 SM-cycles per 128-byte row per SM (6.4 TB/s chip-wide):
//     one reduction per sampled corner (22.75 M per bs=2 encoder layer) is a 455 us floor; shared-memory float
//     atomics are a CAS loop (14.7 cycles per row), so merging must be "owner computes", and native int
//     ATOMS.ADD (0.1 cycles per lane) makes a counting sort cheap;
//   * encoder self-attention is spatially local: the 64 queries of an 8x8-pixel patch put their
//     64*L*P*4 = 4096 corner contributions per head on ~330 distinct rows (scratch/bbox_stats.py).
//
// Block = one head x a tile of 64 queries (host-provided patch order for encoder self-attention), 256 threads.
//   front end   thread (level, query) decodes 4 sampling points (bit-exact geometry of msda_common.cuh);
//               REDUX + shared atomics give each level's bounding box of (h0, w0).  Levels whose box
//               [hmin, hmax+1] x [wmin, wmax+1] fits what is left of the row pool get a window (coarsest
//               level first: smallest boxes); the block copies the windows with 16-byte cp.async,
//               zero-filling rows outside the image, so windowed samples need no corner predicates.
//               Levels that do not fit stay "direct": their samples gather from global memory.
//   sort        the windowed samples are counting-sorted by cell (= pool row of corner (h0,w0)).
//   sorted pass each 4-lane group (8 channels per lane) walks a contiguous chunk of the sorted list holding
//               the current cell's four value rows and four grad_value accumulators in registers: per sample
//               two LDS.128 of grad_out, 16 FFMA2 of dot products, 16 FFMA2 of accumulation; a cell change
//               flushes two (adjacent cell: the other two slide over) or four accumulators with
//               REDG.ADD.F32x4.  grad_sampling_loc / grad_attn_weight are parked in the record slots.
//   direct pass one 4-lane group per query: row gathers from global memory with all of a level's loads in
//               flight, one reduction per corner.
//   write-out   coalesced 16-byte streaming stores of the parked gradients.
// Variants: deterministic (canonical order inside the block, 64-bit fixed-point accumulation across blocks)
// and fused prologue (raw offsets / logits in, their gradients out).
//
// The arithmetic restates models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159 (gradients) and
// :285-288 (geometry); nothing of that file's thread mapping or reductions is used.
#pragma once

#include <climits>

#include "msda_d32.cuh"

namespace msda {

// Queries per block (one 8x8 patch of the host's order); 4 threads per query (one per level slot).  Half patches
// (32 queries, 128 threads, four blocks per SM with a 256-row pool) measured 0.413 ms against 0.344 ms per bs=2 encoder
// layer: 40 % more window cells and staged rows per query outweigh the finer-grained overlap of the blocks' phases.
#ifndef MSDA_WIN_TILEQ
#define MSDA_WIN_TILEQ 64
#endif
constexpr int kWinTileQ = MSDA_WIN_TILEQ;
constexpr int kWinThreads = 4 * kWinTileQ;
// Lanes per lane group of the sorted and direct passes (the group covers the 32 channels of a row):
//   4 lanes x 8 channels: value rows + accumulators = 64 registers per thread -> 128 registers, two blocks per SM,
//                         448-row pool (57 KB fp32);
//   8 lanes x 4 channels: 32 registers of rows + accumulators -> <= 85 registers, THREE blocks per SM with a
//                         320-row pool (41 KB): the per-sample scalar work is replicated over twice the lanes, but
//                         the kernel is bound by latency at 16 resident warps (profiles/r2_*), and 24 hide more of it.
#ifndef MSDA_WIN_G
#define MSDA_WIN_G 4
#endif
constexpr int kWinG = MSDA_WIN_G;
static_assert(kWinG == 4 || kWinG == 8, "lane groups of 4 or 8 lanes");
constexpr int kWinNC2 = 16 / kWinG;   // float2 (channel pairs) of a row held by one lane: 4 | 2
constexpr int kWinNCH = 8 / kWinG;    // 16-byte fp32 chunks of a row held by one lane: 2 | 1
// Rows of the window pool (measured per bs=2 encoder layer, round 1, 4-lane groups: 256 rows 0.464 ms, 320 0.442,
// 384 0.425, 448 0.414, 592 0.434).
#ifndef MSDA_WIN_POOL
#define MSDA_WIN_POOL (MSDA_WIN_G == 4 ? 448 : 320)
#endif
constexpr int kWinPool = MSDA_WIN_POOL;
// Order in which the grid walks the tiles (1: last to first, see the kernel).
#ifndef MSDA_WIN_REVERSE
#define MSDA_WIN_REVERSE 1
#endif

// Index checks for debug builds (MSDA_NVCC_EXTRA=-DMSDA_WIN_CHECKS): compute-sanitizer is not available on
// the GPU pool, so the shared-memory indices of this kernel are asserted by hand; a failed check traps.
#ifdef MSDA_WIN_CHECKS
#define WIN_CHECK(cond) do { if (!(cond)) asm volatile("trap;"); } while (0)
#else
#define WIN_CHECK(cond) do { } while (0)
#endif

#ifdef MSDA_WIN_TIMING
// Debug builds only (scratch/win_timing.py): SM-clock sums over all blocks.  [w] = sorted pass of warp w, [8 + w] =
// direct pass of warp w, [16] = front end, [17] = blocks, [20..26] = front-end sections as stamped by thread 0; read
// and cleared by msda_debug_win_timing().
__device__ unsigned long long g_win_timing[32];
#define WIN_STAMP(k) do { if (t == 0) { const long long n_ = clock64(); atomicAdd(&g_win_timing[k], (unsigned long long)(n_ - t_stamp_dbg)); t_stamp_dbg = n_; } } while (0)
#else
#define WIN_STAMP(k) do { } while (0)
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// 16-byte global->shared copy that bypasses L1 and registers; src_bytes = 0 writes zeros.
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// Pulls the line holding p into L2 (no register, no L1): used one wave of blocks ahead of the demand load.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// How many tiles ahead a block prefetches sampling locations / weights / grad_out (~ one wave of blocks:
// 148 SMs x 2 blocks / 8 heads).
#ifndef MSDA_WIN_PREFETCH_TILES
#define MSDA_WIN_PREFETCH_TILES 48
#endif
constexpr int kWinPrefetchTiles = MSDA_WIN_PREFETCH_TILES;

__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <typename VT, int kL>
struct WinCfg {
  static constexpr int LP = kL * 4;
  static constexpr int NLV = (kL + 3) / 4;            // levels decoded per thread
  static constexpr int REC_STRIDE = LP + 1;           // float4 per query (+1: bank skew)
  static constexpr int ROWB = 32 * (int)sizeof(VT);   // bytes of a value row in the pool
  static constexpr int POOL_ROWS = kWinPool + 2;      // + two all-zero rows (skipped samples, list padding)
  static constexpr int REC_SLOTS = kWinTileQ * REC_STRIDE + 1;  // + one slot for the list's padding entries
  static constexpr int HIST_N = ((kWinPool + kWinThreads - 1) / kWinThreads) * kWinThreads;  // padded for the scan
  static constexpr int SPT = HIST_N / kWinThreads;
  static constexpr int SORTED_N = kWinTileQ * LP + 8; // + padding entries
  static constexpr int PAD_SID = kWinTileQ * LP;      // sample id of a padding entry
  static constexpr int a16(int x) { return (x + 15) / 16 * 16; }
  static constexpr int OFF_POOL = 0;
  static constexpr int OFF_REC = OFF_POOL + POOL_ROWS * ROWB;
  static constexpr int OFF_GO = OFF_REC + REC_SLOTS * 16;                 // grad_out rows of the tile as fp32, + one zero row
  static constexpr int OFF_HIST = OFF_GO + (kWinTileQ + 1) * 128;         // per-cell counts, then exclusive offsets
  static constexpr int OFF_ROWOFF = OFF_HIST + a16((HIST_N + 4) * 4);     // per pool row: element offset in the image, -1 outside
  static constexpr int OFF_SORTED = OFF_ROWOFF + a16(POOL_ROWS * 4);      // sample ids sorted by cell
  static constexpr int OFF_MISC = OFF_SORTED + a16(SORTED_N * 2);         // bb[32], misc[32]
  static constexpr int OFF_INFLAG = OFF_MISC + 256;                       // per (level slot, query): bit i = point i in range
  static constexpr int OFF_STATS = OFF_INFLAG + a16(NLV * 4 * kWinTileQ); // fused prologue: softmax (max, sum) per query
  static constexpr int BWD_SMEM = OFF_STATS + kWinTileQ * 8;
  // deterministic mode: per-warp, per-cell sample counts (16-bit) that make the sort ranks scheduling-independent
  static constexpr int OFF_WCNT = a16(BWD_SMEM);
  static constexpr int BWD_DET_SMEM = OFF_WCNT + (kWinThreads / 32) * HIST_N * 2;
  static_assert(kL <= 8, "per-level state is kept in 8-entry arrays");
  static_assert(kWinPool + 2 < 32768, "two pool rows are packed in one record word");
  static_assert(kWinTileQ * LP + 1 < 65536, "sample ids are stored as 16-bit");
  static_assert((POOL_ROWS * ROWB) % 16 == 0, "pool size keeps the records 16-byte aligned");
};

// Per-level window decision, computed identically by every thread from the block's bounding boxes.
template <int kL>
struct WinAlloc {
  int base[kL];        // first pool row of the level's window, -1: level is gathered from global memory
  int bw[kL], bh[kL];  // window size in rows
  int hm[kL], wm[kL];  // window origin (pixel coordinates, may be -1)
};

template <int kL>
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

// One decoded sampling point held in registers between the phases.
struct WinPoint {
  int h0, w0;
  float lh, lw, a;
  bool in;
};

// Record word of a point.
//   level served from the window: pool row of corner (h0,w0) | pool row of corner (h1,w0) << 16; the
//     (.,w1) corners are the next rows.  A skipped sample points at the pool's two all-zero rows (and
//     carries weight 0).
//   level gathered from global memory: element offset of corner (h0,w0)'s row | 4-bit corner mask
//     (as msda_d32.cuh); skipped -> 0.
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

// Softmax statistics of the L*P logits of one (query, head): max and sum of exp(x - max).
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

// Decodes the 4 points of level l of one (query, head) from the level's raw 48 bytes (already in registers) and
// folds them into the thread's bounding box.  fz.ref_dim != 0: the raw values are offsets / logits (fused prologue).
template <int kLP>
__device__ __forceinline__ void win_decode_loaded(float4 xy01, float4 xy23, float4 aw, const int l, const MsdaLevels& lv,
                                                  WinPoint (&pt)[4], int& hmn, int& hmx, int& wmn, int& wmx,
                                                  const MsdaFused fz, const size_t bq, const float2* stats) {
  constexpr int num_levels = kLP / 4;
  const int H = lv.H[l], W = lv.W[l];
  const bool fma = lv.coord_fma != 0;
  if (fz.ref_dim) {  // fused prologue: raw offsets / logits -> locations / weights
    const float2 st = *stats;
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
// consecutive chunks — which is how the sorted pass and the direct pass distribute a row — then add
// into consecutive accumulators, i.e. one 32-byte sector per four lanes (L2 retires atomics per sector).
__device__ __forceinline__ int win_det_pos(const int c) { return 8 * (c & 3) + (c >> 2); }

// How the lanes of a group cover the 32 channels of a row.  A lane owns kWinNCH 16-byte fp32 chunks (chunk = 4
// channels) of every fp32 row (grad_out, grad_value) and the same channels of the value rows.
//   4-lane groups: chunks cA and cA ^ 4, cA = sj for even groups, sj + 4 for odd ones, so that the two groups of a
//                  quarter-warp hit different bank halves (conflict-free LDS.128) and the four lanes of one reduction
//                  instruction cover 64 contiguous bytes = two whole 32-byte sectors;
//   8-lane groups: chunk sj — a quarter-warp reads / reduces one whole 128-byte row.
// bf16 value rows keep the same channel ownership (their chunk is 8 bytes: LDS.64), so the reductions are full
// sectors for them too.
template <typename VT>
struct WinLane {
  const unsigned char* poolA;  // pool + byte offset of chunk A inside a value row
  const unsigned char* poolB;  // chunk B (4-lane groups only)
  __device__ __forceinline__ WinLane(const unsigned char* pool, const int cA)
      : poolA(pool + cA * 4 * (int)sizeof(VT)), poolB(pool + (cA ^ 4) * 4 * (int)sizeof(VT)) {}
  static __device__ __forceinline__ void put(const float4 t, float2* v) { v[0] = make_float2(t.x, t.y); v[1] = make_float2(t.z, t.w); }
  static __device__ __forceinline__ void put(const uint2 t, float2* v) {
    v[0] = make_float2(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u));
    v[1] = make_float2(__uint_as_float(t.y << 16), __uint_as_float(t.y & 0xffff0000u));
  }
  // channels of chunk A -> v[0], v[1]; chunk B -> v[2], v[3]
  __device__ __forceinline__ void row(const int r, float2 (&v)[kWinNC2]) const {
    if constexpr (sizeof(VT) == 4) {
      put(*reinterpret_cast<const float4*>(poolA + r * 128), v);
      if constexpr (kWinNCH == 2) put(*reinterpret_cast<const float4*>(poolB + r * 128), v + 2);
    } else {
      put(*reinterpret_cast<const uint2*>(poolA + r * 64), v);
      if constexpr (kWinNCH == 2) put(*reinterpret_cast<const uint2*>(poolB + r * 64), v + 2);
    }
  }
  // the same chunks of a value row in global memory (direct pass); p points at chunk A of the row, dB is the
  // element distance to chunk B
  static __device__ __forceinline__ void ldg(const VT* p, const int dB, float2 (&v)[kWinNC2]) {
    if constexpr (sizeof(VT) == 4) {
      put(__ldg(reinterpret_cast<const float4*>(p)), v);
      if constexpr (kWinNCH == 2) put(__ldg(reinterpret_cast<const float4*>(p + dB)), v + 2);
    } else {
      put(__ldg(reinterpret_cast<const uint2*>(p)), v);
      if constexpr (kWinNCH == 2) put(__ldg(reinterpret_cast<const uint2*>(p + dB)), v + 2);
    }
  }
};

// Adds this lane's channels of a partial row into grad_value row `off` (an element offset inside the image; off < 0:
// a pool row outside the image, nothing to add).  Atomic mode: REDG.E.ADD.F32x4 per chunk; deterministic mode:
// 64-bit fixed-point adds.
template <bool kDet>
struct WinRed {
  float* gvA;      // grad_value + image + 4 * cA
  float* gvB;
  long long* g64;  // deterministic: accumulators + image + cA
  int dB64;        // (cA ^ 4) - cA
  float dscale;
  int spread_rows; // knock-out experiments only
  __device__ __forceinline__ void operator()(const int off, const float2 (&acc)[kWinNC2]) const {
#if defined(MSDA_WIN_KNOCKOUT) && (MSDA_WIN_KNOCKOUT & 1)
    if (off == 0x7fffffff) red_add_f4(gvA, acc[0].x, acc[0].y, acc[1].x, acc[1].y);  // never true: keeps the operands alive
    return;
#endif
#if defined(MSDA_WIN_KNOCKOUT) && (MSDA_WIN_KNOCKOUT & 16)
    // same number of reductions, but every block scatters them over the whole image instead of its own (shared, hot) rows
    if (off >= 0) {
      const int rows = spread_rows;
      const int r2 = (int)(((unsigned)(off >> 5) * 2654435761u + blockIdx.x * 40503u) % (unsigned)rows);
      red_add_2xf4_if(gvA, gvB, r2 << 5, acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[kWinNC2 - 2].x, acc[kWinNC2 - 2].y, acc[kWinNC2 - 1].x, acc[kWinNC2 - 1].y);
    }
    return;
#endif
    if constexpr (kDet) {
      if (off >= 0) {
        long long* pa = g64 + off;
        win_det_add(pa, acc[0].x, dscale); win_det_add(pa + 8, acc[0].y, dscale);
        win_det_add(pa + 16, acc[1].x, dscale); win_det_add(pa + 24, acc[1].y, dscale);
        if constexpr (kWinNCH == 2) {
          long long* pb = pa + dB64;
          win_det_add(pb, acc[2].x, dscale); win_det_add(pb + 8, acc[2].y, dscale);
          win_det_add(pb + 16, acc[3].x, dscale); win_det_add(pb + 24, acc[3].y, dscale);
        }
      }
    } else if constexpr (kWinNCH == 2) {
      red_add_2xf4_if(gvA, gvB, off, acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, acc[3].x, acc[3].y);
    } else {
      if (off >= 0) red_add_f4(gvA + off, acc[0].x, acc[0].y, acc[1].x, acc[1].y);
    }
  }
};

// Reduce-scatter of the partial sums of 4 samples over the lanes of a group: afterwards the lane(s) that own sample
// u hold the group's total of v[u].  4-lane groups: lane u owns sample u; 8-lane groups: lanes 2u and 2u + 1.
__device__ __forceinline__ float win_reduce_scatter4(float (&v)[4], const int sj, const unsigned gmask) {
  if constexpr (kWinG == 4) {
    return group_reduce_scatter<4>(v, sj, gmask);
  } else {
    const bool hi = sj & 4;
    float a0 = hi ? v[2] : v[0], a1 = hi ? v[3] : v[1];
    a0 += __shfl_xor_sync(gmask, hi ? v[0] : v[2], 4);
    a1 += __shfl_xor_sync(gmask, hi ? v[1] : v[3], 4);
    const bool hi2 = sj & 2;
    float bsum = hi2 ? a1 : a0;
    bsum += __shfl_xor_sync(gmask, hi2 ? a0 : a1, 2);
    return bsum + __shfl_xor_sync(gmask, bsum, 1);
  }
}

#ifndef MSDA_WIN_BWD_MINBLOCKS
#define MSDA_WIN_BWD_MINBLOCKS (MSDA_WIN_G == 4 ? 2 : 3)
#endif

// One block = one tile x one head.  kDet: deterministic grad_value (canonical order inside the block,
// order-independent fixed-point accumulation across blocks; see msda_launch_win.cu).
template <typename VT, int kL, int kM, bool kDet, bool kFused>
__global__ void __launch_bounds__(kWinThreads, MSDA_WIN_BWD_MINBLOCKS)
msda_bwd_d32_win_kernel(const WinBwdArgs ar, const __grid_constant__ MsdaLevels lv) {
  using Cfg = WinCfg<VT, kL>;
  constexpr int LP = Cfg::LP, NLV = Cfg::NLV, ROWB = Cfg::ROWB, RS = Cfg::REC_STRIDE;
  extern __shared__ __align__(128) unsigned char smraw[];
  unsigned char* pool = smraw + Cfg::OFF_POOL;
  float4* rec = reinterpret_cast<float4*>(smraw + Cfg::OFF_REC);
  float* go_s = reinterpret_cast<float*>(smraw + Cfg::OFF_GO);
  int* hist = reinterpret_cast<int*>(smraw + Cfg::OFF_HIST);
  int* rowoff = reinterpret_cast<int*>(smraw + Cfg::OFF_ROWOFF);
  unsigned short* sorted = reinterpret_cast<unsigned short*>(smraw + Cfg::OFF_SORTED);
  int* bb = reinterpret_cast<int*>(smraw + Cfg::OFF_MISC);  // [0,8) hmin [8,16) hmax [16,24) wmin [24,32) wmax
  int* misc = bb + 32;                                      // [0,8) warp totals [16] total [20,28) window base per level
  unsigned char* inflag = smraw + Cfg::OFF_INFLAG;
  float2* stats = reinterpret_cast<float2*>(smraw + Cfg::OFF_STATS);
  unsigned short* wcnt = reinterpret_cast<unsigned short*>(smraw + Cfg::OFF_WCNT);  // deterministic mode only

  const int M = kM ? kM : ar.M, Lq = ar.Lq, S = ar.S, M32 = M * 32;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // Tiles are handed out last to first: the host's patch order ends with the coarse levels' queries, whose tiles are
  // the expensive ones (their fine-level samples do not fit the pool and take the direct pass) — started first, they
  // do not end up alone in the kernel's tail.
  const int m = blockIdx.x % M, b = blockIdx.y;
  const int tile = MSDA_WIN_REVERSE ? (int)(gridDim.x / M) - 1 - (int)(blockIdx.x / M) : (int)(blockIdx.x / M);
  constexpr int kPfStep = MSDA_WIN_REVERSE ? -kWinPrefetchTiles : kWinPrefetchTiles;  // the tile one wave of blocks later
  const int* order = ar.order;
  const int order_len = ar.order_len;
  const VT* grad_out = static_cast<const VT*>(ar.grad_out);
  const size_t img = (size_t)b * S * M32;
  const VT* value_img = static_cast<const VT*>(ar.value) + img;
  const MsdaFused fz = kFused ? ar.fz : MsdaFused{nullptr, 0};

  // thread <-> (level slot, query): the level is warp-uniform
  const int ql = t & (kWinTileQ - 1), slot = t / kWinTileQ;
  auto tile_query = [&](const int tl, const int k) {
    const int os = tl * kWinTileQ + k;
    return (os >= 0 && os < order_len) ? (order ? order[os] : os) : -1;
  };
  // every order entry this thread needs is requested before the first load that depends on one (one L2 round trip
  // instead of three back to back)
  const int q = tile_query(tile, ql);
  const int qpf = tile_query(tile + kPfStep, ql);
  int gq_[2];
  if constexpr (sizeof(VT) == 4) {
    gq_[0] = tile_query(tile, t >> 3);
    gq_[1] = tile_query(tile, (t + kWinThreads) >> 3);
  } else {
    gq_[0] = gq_[1] = tile_query(tile, t >> 2);
  }
  const size_t bq = (size_t)b * Lq + (q >= 0 ? q : 0);
  const size_t qm = bq * M + m;

#ifdef MSDA_WIN_TIMING
  const long long t_start_dbg = clock64();
  long long t_stamp_dbg = t_start_dbg;
#endif
  // ---- phase 0: every global load of the front end goes out first -----------------------------------
  float4 rxy01[NLV], rxy23[NLV], raw[NLV];
#pragma unroll
  for (int li = 0; li < NLV; ++li) {
    const int l = slot + 4 * li;
    rxy01[li] = rxy23[li] = raw[li] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < kL && q >= 0) {
      const float* lp = ar.loc + (qm * LP + l * 4) * 2;
      rxy01[li] = ld_stream_f4(lp);
      rxy23[li] = ld_stream_f4(lp + 4);
      raw[li] = ld_stream_f4(ar.attw + qm * LP + l * 4);
    }
  }
  // grad_out rows of the tile -> go_s (fp32).  fp32: straight into shared memory (cp.async, zero-filled for the
  // tile's empty slots); bf16: through registers, converted and stored after the decode.
  float go_reg[8];
  if constexpr (sizeof(VT) == 4) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int i = t + it * kWinThreads, gql = i >> 3, gj = i & 7;
      const int gq = gq_[it];
      cp_async16(smem_u32(go_s + gql * 32 + gj * 4),
                 grad_out + (((size_t)b * Lq + (gq >= 0 ? gq : 0)) * M + m) * 32 + gj * 4, gq >= 0 ? 16 : 0);
    }
  } else {
    const int gj = t & 3;
    const int gq = gq_[0];
#pragma unroll
    for (int c = 0; c < 8; ++c) go_reg[c] = 0.f;
    if (gq >= 0) RowTraits<__nv_bfloat16>::load_stream(grad_out + (((size_t)b * Lq + gq) * M + m) * 32 + gj * 8, go_reg);
  }
  {  // a later tile's inputs: HBM -> L2 now, so that its front end sees L2 latency
    if (qpf >= 0) {
      const size_t qm_pf = ((size_t)b * Lq + qpf) * M + m;
#pragma unroll
      for (int li = 0; li < NLV; ++li) {
        const int l = slot + 4 * li;
        if (l < kL) {
          prefetch_l2(ar.loc + (qm_pf * LP + l * 4) * 2);
          prefetch_l2(ar.attw + qm_pf * LP + l * 4);
        }
      }
      if (slot == 0) prefetch_l2(grad_out + qm_pf * 32);
    }
  }
  // fused prologue: the level-slot-0 thread of each query takes the softmax statistics of its L*P logits
  if (kFused && slot == 0 && q >= 0) stats[ql] = win_softmax_stats<LP>(ar.attw + qm * LP);
  if (t < 32) bb[t] = (t & 8) ? INT_MIN : INT_MAX;
  if (t >= 32 && t < 32 + 2 * ROWB / 16)  // the pool's two all-zero rows
    reinterpret_cast<uint4*>(pool + kWinPool * ROWB)[t - 32] = make_uint4(0u, 0u, 0u, 0u);
  if (t >= 64 && t < 72) reinterpret_cast<float4*>(go_s + kWinTileQ * 32)[t - 64] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t == 72) rec[kWinTileQ * RS] = make_float4(__int_as_float(kWinPool | (kWinPool << 16)), 0.f, 0.f, 0.f);
  if (t == 73 || t == 74) rowoff[kWinPool + t - 73] = -1;
#pragma unroll
  for (int k = 0; k < Cfg::SPT; ++k) hist[t + k * kWinThreads] = 0;
  if (kDet) {
    for (int i = t; i < (kWinThreads / 32) * Cfg::HIST_N * 2 / 16; i += kWinThreads)
      reinterpret_cast<uint4*>(wcnt)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  WIN_STAMP(20);  // loads issued, init, barrier

  // ---- phase 1: decode, bounding boxes ---------------------------------------------------------------
  WinPoint pts[NLV][4];
#pragma unroll
  for (int li = 0; li < NLV; ++li) {
    const int l = slot + 4 * li;  // warp-uniform
    int hmn = INT_MAX, hmx = INT_MIN, wmn = INT_MAX, wmx = INT_MIN;
#pragma unroll
    for (int i = 0; i < 4; ++i) pts[li][i] = WinPoint{0, 0, 0.f, 0.f, 0.f, false};
    if (l < kL) {
      if (q >= 0) win_decode_loaded<LP>(rxy01[li], rxy23[li], raw[li], l, lv, pts[li], hmn, hmx, wmn, wmx, fz, bq, stats + ql);
      hmn = __reduce_min_sync(0xffffffffu, hmn); hmx = __reduce_max_sync(0xffffffffu, hmx);
      wmn = __reduce_min_sync(0xffffffffu, wmn); wmx = __reduce_max_sync(0xffffffffu, wmx);
      if (lane == 0 && hmn <= hmx) {
        atomicMin(&bb[l], hmn); atomicMax(&bb[8 + l], hmx);
        atomicMin(&bb[16 + l], wmn); atomicMax(&bb[24 + l], wmx);
      }
    }
  }
  __syncthreads();
  WIN_STAMP(21);  // decode (waits for the loads), bounding boxes, barrier

  // ---- phase 2: windows (cp.async left in flight), records, per-cell counts ---------------------------
  int rank[NLV][4];
  {
    WinAlloc<kL> wa;
    win_allocate<kL>(bb, wa);
    {  // staging: the rows of a level's window are dealt round-robin to groups of G lanes (16 bytes per lane)
      constexpr int G = ROWB / 16, EPL = 16 / (int)sizeof(VT);
      const int r0 = t / G, jj = t % G;
      const unsigned pool_s = smem_u32(pool) + jj * 16;
      const VT* src_j = value_img + jj * EPL;
#pragma unroll
      for (int l = kL - 1; l >= 0; --l) {
        if (wa.base[l] < 0) continue;  // block-uniform
        const int bw = wa.bw[l], n = wa.bh[l] * bw, H = lv.H[l], W = lv.W[l];
        // r / bw through a reciprocal: (r + 0.5) / bw is at least 0.5 / bw >= 1e-3 away from an integer and the
        // approximation error is below 1e-4 for r < 2^15, so the truncation is exact
        const float inv_bw = __fdividef(1.f, (float)bw);
        const int off00 = ((lv.start[l] + wa.hm[l] * W + wa.wm[l]) * M + m) * 32;
#if defined(MSDA_WIN_KNOCKOUT) && (MSDA_WIN_KNOCKOUT & 32)
        if (t >= 0) continue;  // knock-out: no staging loop at all (use together with bit 0: row offsets stay unset)
#endif
        for (int r = r0; r < n; r += kWinThreads / G) {
          const int rh = (int)(((float)r + 0.5f) * inv_bw), rw = r - rh * bw;
          const bool inb = (unsigned)(wa.hm[l] + rh) < (unsigned)H && (unsigned)(wa.wm[l] + rw) < (unsigned)W;
          const int off = off00 + (rh * W + rw) * M32;
          WIN_CHECK(rw >= 0 && rw < bw && wa.base[l] + r < kWinPool);
#if defined(MSDA_WIN_KNOCKOUT) && (MSDA_WIN_KNOCKOUT & 2)
          cp_async16(pool_s + (unsigned)((wa.base[l] + r) * ROWB), src_j, 0);  // zero-fill only: no global read
#else
          cp_async16(pool_s + (unsigned)((wa.base[l] + r) * ROWB), src_j + (inb ? off : 0), inb ? 16 : 0);
#endif
          if (jj == 0) rowoff[wa.base[l] + r] = inb ? off : -1;
        }
      }
    }
#pragma unroll
    for (int li = 0; li < NLV; ++li) {
      const int l = slot + 4 * li;
#pragma unroll
      for (int i = 0; i < 4; ++i) rank[li][i] = -1;
      if (l < kL) {
        int base = -1, bw = 0, hm = 0, wm = 0;
#pragma unroll
        for (int ll = 0; ll < kL; ++ll)
          if (ll == l) { base = wa.base[ll]; bw = wa.bw[ll]; hm = wa.hm[ll]; wm = wa.wm[ll]; }
        const int H = lv.H[l], W = lv.W[l], st = lv.start[l];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int code = win_record_code(pts[li][i], base, bw, hm, wm, H, W, st, m, M);
          WIN_CHECK(base < 0 || ((code & 0xffff) + 1 <= kWinPool + 1 && (code >> 16) + 1 <= kWinPool + 1));
          WIN_CHECK(base < 0 || !pts[li][i].in || ((code >> 16) + 1 < kWinPool && (code & 0xffff) >= base));
          const bool part = base >= 0 && pts[li][i].in;  // takes part in the sort
          if (!kDet) {
            if (part) rank[li][i] = atomicAdd(&hist[code & 0xffff], 1);  // arrival order inside the cell
          } else {
            // scheduling-independent rank: lanes of a warp that hit the same cell are ranked in lane order on
            // top of what the warp counted for that cell in earlier rounds; the warps' counts are turned into
            // exclusive bases after the block barrier
            const int cell = code & 0xffff;
            const unsigned peers = __match_any_sync(0xffffffffu, part ? (unsigned)cell : (0xffff0000u | (unsigned)lane));
            const int leader = __ffs(peers) - 1;
            unsigned short* cnt = wcnt + warp * Cfg::HIST_N + cell;
            int prev = 0;
            if (part && lane == leader) { prev = *cnt; *cnt = (unsigned short)(prev + __popc(peers)); }
            prev = __shfl_sync(0xffffffffu, prev, leader);
            if (part) rank[li][i] = prev + __popc(peers & ((1u << lane) - 1u));
            __syncwarp();
          }
          rec[ql * RS + l * 4 + i] =
              make_float4(__int_as_float(code), pts[li][i].lh, pts[li][i].lw, pts[li][i].in ? pts[li][i].a : 0.f);
        }
      }
      inflag[(li * 4 + slot) * kWinTileQ + ql] =
          (unsigned char)((pts[li][0].in ? 1 : 0) | (pts[li][1].in ? 2 : 0) | (pts[li][2].in ? 4 : 0) | (pts[li][3].in ? 8 : 0));
    }
    if (t < kL) {
#pragma unroll
      for (int l = 0; l < kL; ++l)
        if (l == t) misc[20 + l] = wa.base[l];
    }
  }
  if constexpr (sizeof(VT) != 4) {
    float* d = go_s + (t >> 2) * 32 + (t & 3) * 8;
    *reinterpret_cast<float4*>(d) = make_float4(go_reg[0], go_reg[1], go_reg[2], go_reg[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(go_reg[4], go_reg[5], go_reg[6], go_reg[7]);
  }
  __syncthreads();  // counts complete, records visible
  WIN_STAMP(22);  // allocation, staging issue, records, counts, barrier

  // ---- phase 3: counting sort by cell ---------------------------------------------------------------
  if (kDet) {
    // per cell: the warps' counts -> exclusive bases over the warps, their sum -> hist
    for (int c = t; c < Cfg::HIST_N; c += kWinThreads) {
      int run = 0;
#pragma unroll
      for (int w = 0; w < kWinThreads / 32; ++w) {
        const int n = wcnt[w * Cfg::HIST_N + c];
        wcnt[w * Cfg::HIST_N + c] = (unsigned short)run;
        run += n;
      }
      hist[c] = run;
    }
    __syncthreads();
  }
  {  // exclusive scan of the per-cell counts, in place
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
#pragma unroll
    for (int w = 0; w < kWinThreads / 32; ++w)
      if (w < warp) run += misc[w];
#pragma unroll
    for (int k = 0; k < Cfg::SPT; ++k) { hist[t * Cfg::SPT + k] = run; run += v[k]; }
    if (t == kWinThreads - 1) misc[16] = run;
    __syncthreads();
  }
  WIN_STAMP(23);  // scan (two barriers)
  // sample ids to their sorted positions; the list is padded with entries that point at the all-zero rows
#pragma unroll
  for (int li = 0; li < NLV; ++li) {
    const int l = slot + 4 * li;
    if (l < kL) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (rank[li][i] >= 0) {
          const int cell = __float_as_int(rec[ql * RS + l * 4 + i].x) & 0xffff;
          const int wbase = kDet ? (int)wcnt[warp * Cfg::HIST_N + cell] : 0;
          WIN_CHECK(hist[cell] + wbase + rank[li][i] >= 0 && hist[cell] + wbase + rank[li][i] < misc[16]);
          sorted[hist[cell] + wbase + rank[li][i]] = (unsigned short)(ql * LP + l * 4 + i);
        }
    }
  }
  if (t < 8) sorted[misc[16] + t] = (unsigned short)Cfg::PAD_SID;
  WIN_STAMP(24);  // placement
  cp_async_wait_all();
  WIN_STAMP(25);  // wait for the windows (thread 0's own copies)
  __syncthreads();  // windows, sorted list, grad_out rows are in shared memory
  WIN_STAMP(26);  // final barrier of the front end

  // the front end writes nothing to global memory; everything from here on may reduce into grad_value, which the
  // preceding kernel on the stream zero-fills when this kernel was allowed to start early (programmatic dependent
  // launch, msda_capi.cu); otherwise the wait returns at once
  if (!kDet) asm volatile("griddepcontrol.wait;" ::: "memory");

  // lane roles of the sorted and direct passes: 32 / kWinG groups of kWinG lanes per warp
  constexpr int GPW = 32 / kWinG;                 // lane groups per warp
  const int sg = lane / kWinG, sj = lane % kWinG;
  const int cA = kWinG == 4 ? sj + ((sg & 1) ? 4 : 0) : sj, cB = cA ^ 4;  // this lane's 16-byte chunk(s) of an fp32 row
  const unsigned gmask = (kWinG == 4 ? 0xfu : 0xffu) << (sg * kWinG);
  const unsigned char* goA = reinterpret_cast<const unsigned char*>(go_s) + cA * 16;
  const unsigned char* goB = reinterpret_cast<const unsigned char*>(go_s) + cB * 16;
  // grad_out row of tile query `sq`, this lane's channels as pairs
  auto load_go = [&](const int sq, float2 (&go)[kWinNC2]) {
    const float4 a4 = *reinterpret_cast<const float4*>(goA + sq * 128);
    go[0] = make_float2(a4.x, a4.y); go[1] = make_float2(a4.z, a4.w);
    if constexpr (kWinNCH == 2) {
      const float4 b4 = *reinterpret_cast<const float4*>(goB + sq * 128);
      go[2] = make_float2(b4.x, b4.y); go[3] = make_float2(b4.z, b4.w);
    }
  };
  // which lane parks sample u of a step of 4, and which sample this lane parks
  const int my_u = kWinG == 4 ? sj : (sj >> 1);
  const bool parks = kWinG == 4 ? true : (sj & 1) == 0;
  WinRed<kDet> red;
  red.gvA = ar.grad_value + img + cA * 4;
  red.gvB = ar.grad_value + img + cB * 4;
  red.g64 = kDet ? ar.gv64 + img + cA : nullptr;
  red.dB64 = cB - cA;
  red.dscale = kDet ? win_det_scale(ar.maxbits, Lq, LP) : 0.f;
  red.spread_rows = S * M;
  const float2 zero2 = make_float2(0.f, 0.f);
  auto sid_of = [](const uint2 pk, const int u) {
    return (int)(u == 0 ? pk.x & 0xffffu : u == 1 ? pk.x >> 16 : u == 2 ? pk.y & 0xffffu : pk.y >> 16);
  };
  auto rec_slot = [](const int sid) { return sid + sid / LP; };  // query * (LP + 1) + point

#ifdef MSDA_WIN_TIMING
  long long t_dbg = clock64();
  if (t == 0) { atomicAdd(&g_win_timing[16], (unsigned long long)(t_dbg - t_start_dbg)); atomicAdd(&g_win_timing[17], 1ull); }
#endif
  // ---- phase 4: sorted pass ----------------------------------------------------------------------------
  {
    const WinLane<VT> wl(pool, cA);
    const int total = misc[16], totp = (total + 3) & ~3;   // padded to whole steps of 4 samples
    constexpr int SNG = kWinThreads / kWinG;               // lane groups per block
    const int chunk = ((totp / 4 + SNG - 1) / SNG) * 4;
    const int gi = warp * GPW + sg;
#if defined(MSDA_WIN_KNOCKOUT) && (MSDA_WIN_KNOCKOUT & 4)
    const int i0 = 0, i1 = 0;  // knock-out: no sorted pass
#else
    const int i0 = min(totp, gi * chunk), i1 = min(totp, i0 + chunk);
#endif
    float2 V00[kWinNC2], V01[kWinNC2], V10[kWinNC2], V11[kWinNC2], A0[kWinNC2], A1[kWinNC2], B0[kWinNC2], B1[kWinNC2];
#pragma unroll
    for (int c = 0; c < kWinNC2; ++c) V00[c] = V01[c] = V10[c] = V11[c] = A0[c] = A1[c] = B0[c] = B1[c] = zero2;
    int cur0 = kWinPool, cur1 = kWinPool;  // the all-zero rows: nothing to flush
    auto flush = [&](const int row, const float2 (&acc)[kWinNC2]) {
      WIN_CHECK(row >= 0 && row < kWinPool + 2);
      const int off = rowoff[row];
      WIN_CHECK(off < 0 || (off % 32 == 0 && off / 32 < S * M));
      red(off, acc);
    };
    uint2 pk_next = *reinterpret_cast<const uint2*>(sorted + i0);  // i0 is a multiple of 4
    float4 rnext = rec[rec_slot(sid_of(pk_next, 0))];
    for (int ib = i0; ib < i1; ib += 4) {
      const uint2 pk = pk_next;
      pk_next = *reinterpret_cast<const uint2*>(sorted + ib + 4);  // the list is padded: always readable
      float pgx[4], pgy[4], pga[4], aws[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sq = sid_of(pk, u) / LP;
        const float4 r = rnext;
        // grad_out row of the sample's query; the next sample's record is fetched one step ahead
        float2 go[kWinNC2];
        load_go(sq, go);
        rnext = rec[rec_slot(u < 3 ? sid_of(pk, u + 1) : sid_of(pk_next, 0))];
        const int code = __float_as_int(r.x);
        const int row0 = code & 0xffff, row1 = code >> 16;
        WIN_CHECK(sid_of(pk, u) <= Cfg::PAD_SID && row1 + 1 < kWinPool + 2 && row1 >= row0);
        if (row0 != cur0) {
          const bool adj = (row0 == cur0 + 1);
          flush(cur0, A0);
          flush(cur1, B0);
          if (!adj) {
            flush(cur0 + 1, A1);
            flush(cur1 + 1, B1);
          }
#pragma unroll
          for (int c = 0; c < kWinNC2; ++c) {
            A0[c] = adj ? A1[c] : zero2;
            B0[c] = adj ? B1[c] : zero2;
            A1[c] = zero2;
            B1[c] = zero2;
          }
          wl.row(row0, V00);
          wl.row(row0 + 1, V01);
          wl.row(row1, V10);
          wl.row(row1 + 1, V11);
          cur0 = row0; cur1 = row1;
        }
        const float lh = r.y, lw = r.z, a = r.w;
        aws[u] = a;
        const float hh = 1.f - lh, hw = 1.f - lw;
        // grad_value partial sums (cuh:125,134,143,152): independent of the value rows, so they cover
        // the latency of a window reload
        const float c00 = hh * hw, c01 = hh * lw, c10 = lh * hw, c11 = lh * lw;
        const float w00 = c00 * a, w01 = c01 * a, w10 = c10 * a, w11 = c11 * a;
        const float2 w00p = make_float2(w00, w00), w01p = make_float2(w01, w01), w10p = make_float2(w10, w10),
                     w11p = make_float2(w11, w11);
#pragma unroll
        for (int c = 0; c < kWinNC2; ++c) {
          A0[c] = ffma2(w00p, go[c], A0[c]); A1[c] = ffma2(w01p, go[c], A1[c]);
          B0[c] = ffma2(w10p, go[c], B0[c]); B1[c] = ffma2(w11p, go[c], B1[c]);
        }
        float2 e00 = zero2, e01 = zero2, e10 = zero2, e11 = zero2;
#pragma unroll
        for (int c = 0; c < kWinNC2; ++c) {
          e00 = ffma2(go[c], V00[c], e00); e01 = ffma2(go[c], V01[c], e01);
          e10 = ffma2(go[c], V10[c], e10); e11 = ffma2(go[c], V11[c], e11);
        }
#if defined(MSDA_WIN_EXTRA_FMA)
        // issue-sensitivity experiment: MSDA_WIN_EXTRA_FMA independent FFMA2 per sample that feed nothing that matters
#pragma unroll
        for (int x = 0; x < MSDA_WIN_EXTRA_FMA; ++x) e00 = ffma2(make_float2(1e-30f, 1e-30f), go[x % kWinNC2], e00);
#endif
        const float d00 = e00.x + e00.y, d01 = e01.x + e01.y, d10 = e10.x + e10.y, d11 = e11.x + e11.y;
        // this lane's channels of grad_attn_weight (cuh:156) and of grad_sampling_loc (cuh:157-158) before the
        // factors W_l, H_l, which the write-out applies
        pga[u] = fmaf(c00, d00, fmaf(c01, d01, fmaf(c10, d10, c11 * d11)));
        pgx[u] = a * fmaf(hh, d01 - d00, lh * (d11 - d10));
        pgy[u] = a * fmaf(hw, d10 - d00, lw * (d11 - d01));
      }
      const float gx = win_reduce_scatter4(pgx, sj, gmask);
      const float gy = win_reduce_scatter4(pgy, sj, gmask);
      const float ga = win_reduce_scatter4(pga, sj, gmask);
      // the lane that owns sample ib + my_u parks its gradients in the sample's record slot (.w keeps the sample's
      // weight: the fused prologue's softmax backward needs it); padding entries have no slot of their own
      const int sid = my_u == 0 ? sid_of(pk, 0) : my_u == 1 ? sid_of(pk, 1) : my_u == 2 ? sid_of(pk, 2) : sid_of(pk, 3);
      if (parks && sid < Cfg::PAD_SID)
        rec[rec_slot(sid)] = make_float4(gx, gy, ga, my_u == 0 ? aws[0] : my_u == 1 ? aws[1] : my_u == 2 ? aws[2] : aws[3]);
    }
    flush(cur0, A0);
    flush(cur1, B0);
    flush(cur0 + 1, A1);
    flush(cur1 + 1, B1);
  }

#ifdef MSDA_WIN_TIMING
  __syncwarp();
  if (lane == 0) { const long long n = clock64(); atomicAdd(&g_win_timing[warp], (unsigned long long)(n - t_dbg)); t_dbg = n; }
#endif
  // ---- phase 5: direct pass — levels that did not get a window; one lane group per query ------------------
  {
    int lbase[kL];
    bool all_win = true;
#pragma unroll
    for (int l = 0; l < kL; ++l) { lbase[l] = misc[20 + l]; all_win = all_win && lbase[l] >= 0; }
#if defined(MSDA_WIN_KNOCKOUT) && (MSDA_WIN_KNOCKOUT & 8)
    all_win = true;  // knock-out: no direct pass
#endif
    if (!all_win) {
      const VT* value_A = value_img + cA * 4;
      const int dB = (cB - cA) * 4;
#pragma unroll 1
      for (int dql = warp * GPW + sg; dql < kWinTileQ; dql += (kWinThreads / 32) * GPW) {
        float2 go[kWinNC2];
        load_go(dql, go);
#pragma unroll
        for (int l = 0; l < kL; ++l) {
          if (lbase[l] >= 0) continue;  // block-uniform
          const int o_line = lv.W[l] * M32;
          float4 r[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) r[i] = rec[dql * RS + l * 4 + i];
          float pgx[4], pgy[4], pga[4];
          // two points at a time: their eight row loads go out together (the registers of the sorted pass's rows and
          // accumulators are free here), so a level costs two L2 round trips per query instead of four
#pragma unroll
          for (int ip = 0; ip < 4; ip += 2) {
            float2 v[2][4][kWinNC2];
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
              const int code = __float_as_int(r[ip + ii].x);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int c = 0; c < kWinNC2; ++c) v[ii][k][c] = zero2;
                if (code & (1 << k))
                  WinLane<VT>::ldg(value_A + (code & ~31) + ((k & 2) ? o_line : 0) + ((k & 1) ? M32 : 0), dB, v[ii][k]);
              }
            }
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
              const int i = ip + ii;
              const int code = __float_as_int(r[i].x);
              const float lh = r[i].y, lw = r[i].z, a = r[i].w;
              const float hh = 1.f - lh, hw = 1.f - lw;
              const float a_hh = a * hh, a_lh = a * lh;
              float d[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 e = zero2;
#pragma unroll
                for (int c = 0; c < kWinNC2; ++c) e = ffma2(go[c], v[ii][k][c], e);
                d[k] = e.x + e.y;
                if (code & (1 << k)) {
                  const float tt = ((k & 2) ? a_lh : a_hh) * ((k & 1) ? lw : hw);
                  float2 acc[kWinNC2];
#pragma unroll
                  for (int c = 0; c < kWinNC2; ++c) acc[c] = make_float2(tt * go[c].x, tt * go[c].y);
                  red((code & ~31) + ((k & 2) ? o_line : 0) + ((k & 1) ? M32 : 0), acc);
                }
              }
              pga[i] = hh * (hw * d[0] + lw * d[1]) + lh * (hw * d[2] + lw * d[3]);
              pgx[i] = a * (hh * (d[1] - d[0]) + lh * (d[3] - d[2]));
              pgy[i] = a * (hw * (d[2] - d[0]) + lw * (d[3] - d[1]));
            }
          }
          const float gx = win_reduce_scatter4(pgx, sj, gmask);
          const float gy = win_reduce_scatter4(pgy, sj, gmask);
          const float ga = win_reduce_scatter4(pga, sj, gmask);
          if (parks)
            rec[dql * RS + l * 4 + my_u] = make_float4(gx, gy, ga, my_u == 0 ? r[0].w : my_u == 1 ? r[1].w : my_u == 2 ? r[2].w : r[3].w);
        }
      }
    }
  }
#ifdef MSDA_WIN_TIMING
  __syncwarp();
  if (lane == 0) atomicAdd(&g_win_timing[8 + warp], (unsigned long long)(clock64() - t_dbg));
#endif
  __syncthreads();  // every sample's gradients are parked in its record slot

  // ---- phase 6: write-out: thread <-> (level slot, query) as in the decode -------------------------------
  if (q >= 0) {
#pragma unroll
    for (int li = 0; li < NLV; ++li) {
      const int l = slot + 4 * li;
      if (l < kL) {
        const int fl = inflag[(li * 4 + slot) * kWinTileQ + ql];
        const float Wf = (float)lv.W[l], Hf = (float)lv.H[l];
        float4 r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r[i] = rec[ql * RS + l * 4 + i];
          r[i].x *= Wf;  // (cuh:157-158)
          r[i].y *= Hf;
          if (!(fl & (1 << i))) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // skipped sample: slot still holds its record
        }
        if (kFused) {
          // fused prologue: gradients of the raw offsets and logits.  softmax backward needs
          // dot = sum_j a_j * dL/da_j over all L*P samples of (query, head): the other levels' dL/da sit in
          // their record slots, the weights are recomputed from the logits and the saved (max, sum)
          float dot = 0.f;
          for (int sl = 0; sl < NLV * 4; ++sl) {
            const int l2 = (sl & 3) + 4 * (sl >> 2);
            if (l2 < kL) {
              const int f2 = inflag[sl * kWinTileQ + ql];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 o = rec[ql * RS + l2 * 4 + i];  // (gx, gy, dL/da, a) of a processed sample
                if (f2 & (1 << i)) dot = fmaf(o.w, o.z, dot);
              }
            }
          }
          const float2 st = stats[ql];
          const float4 x = ld_stream_f4(ar.attw + qm * LP + l * 4);
          const float xs[4] = {x.x, x.y, x.z, x.w};
          const float* rp = ar.fz.ref + (bq * kL + l) * ar.fz.ref_dim;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // a skipped sample has no result slot: its weight is recomputed (its dL/da is 0, not its dL/dlogit)
            const float a = (fl & (1 << i)) ? rec[ql * RS + l * 4 + i].w : __expf(xs[i] - st.x) / st.y;
            // d loc / d offset, in autograd's operation order: g / W  |  ((g * 0.5) * wh) / P
            const float gx = ar.fz.ref_dim == 2 ? r[i].x / Wf : r[i].x * 0.5f * rp[2] / 4.f;
            const float gy = ar.fz.ref_dim == 2 ? r[i].y / Hf : r[i].y * 0.5f * rp[3] / 4.f;
            r[i] = make_float4(gx, gy, a * (r[i].z - dot), 0.f);
          }
        }
        float* gl = ar.grad_loc + (qm * LP + l * 4) * 2;
        st_stream_f4(gl, make_float4(r[0].x, r[0].y, r[1].x, r[1].y));
        st_stream_f4(gl + 4, make_float4(r[2].x, r[2].y, r[3].x, r[3].y));
        st_stream_f4(ar.grad_attw + qm * LP + l * 4, make_float4(r[0].z, r[1].z, r[2].z, r[3].z));
      }
    }
  }
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

}  // namespace msda
