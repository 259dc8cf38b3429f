// msda_det.cuh — deterministic grad_value ("sort-by-corner" mode, MSDA_FLAG_DETERMINISTIC).
//
// The atomic backward adds the 4 * N*Lq*M*L*P corner contributions to grad_value in
// whatever order the hardware serves them, so fp32 results differ run to run (the
// reference has the same property, cuh:125-152).  This mode fixes the order:
//
//   1. count   one thread per sample: bump an int32 counter per destination row
//              (b, token, head) for each in-bounds corner            (integer atomics)
//   2. scan    exclusive prefix sum of the counters -> segment offsets
//   3. fill    one thread per sample: append the contribution id (sample*4 + corner)
//              to its destination's segment                (slot order is arbitrary)
//   4. reduce  one warp per destination row: sort the segment's ids ascending (warp
//              bitonic network, in shared memory; in place in the workspace for very
//              long segments), then accumulate coef * grad_out[q,m,:] in that order and
//              store the row once.  Rows without contributions are written as zeros, so
//              no separate zero-fill of grad_value is needed.
//
// Contribution ids are unique, so the ascending order is canonical: the result depends
// only on the inputs, never on scheduling.  grad_sampling_loc / grad_attn_weight come
// from the regular backward kernels with the scatter disabled (they are already
// deterministic: fixed-shape shuffle reductions).
#pragma once

#include <cuda_bf16.h>

#include "msda_common.cuh"
#include "msda_generic.cuh"
#include "msda_host.h"

namespace msda {

constexpr int kDetSortCap = 1024;  // ids per warp sorted in shared memory

struct DetLayout {
  size_t counts_off, offsets_off, ids_off, total;
  long long n_dst, n_sample;
};

inline DetLayout det_layout(int batch, int spatial_size, int num_heads, int num_levels,
                            int num_query, int num_point) {
  DetLayout L;
  L.n_dst = (long long)batch * spatial_size * num_heads;
  L.n_sample = (long long)batch * num_query * num_heads * num_levels * num_point;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  L.counts_off = 0;
  L.offsets_off = up((size_t)(L.n_dst + 1) * 4);
  L.ids_off = L.offsets_off + up((size_t)(L.n_dst + 1) * 4);
  L.total = L.ids_off + up((size_t)L.n_sample * 4 * 4);
  return L;
}

inline size_t deterministic_workspace_bytes(int batch, int spatial_size, int num_heads, int channels,
                                            int num_levels, int num_query, int num_point) {
  (void)channels;
  if (batch < 1 || spatial_size < 1 || num_heads < 1 || num_levels < 1 || num_query < 1 || num_point < 1)
    return 256;
  return det_layout(batch, spatial_size, num_heads, num_levels, num_query, num_point).total;
}

// kFill=false: count; kFill=true: append ids using `counts` as per-destination cursors.
template <bool kFill>
__global__ void __launch_bounds__(256)
msda_det_bin_kernel(const float* __restrict__ loc, int* __restrict__ counts,
                    const unsigned* __restrict__ offsets, unsigned* __restrict__ ids,
                    const __grid_constant__ MsdaLevels lv, const MsdaDims d, const long long n_sample) {
  const int LP = d.num_levels * d.num_point;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_sample;
       s += (long long)gridDim.x * blockDim.x) {
    const int lp = (int)(s % LP);
    const int l = lp / d.num_point;
    const long long qm = s / LP;
    const int m = (int)(qm % d.num_heads);
    const int b = (int)(qm / ((long long)d.num_heads * d.num_query));
    const float2 xy = *reinterpret_cast<const float2*>(loc + s * 2);
    int tok[4];
    float lh, lw;
    msda_sample_geom(xy.x, xy.y, lv.H[l], lv.W[l], lv.start[l], tok, lh, lw, lv.coord_fma != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (tok[k] < 0) continue;
      const long long dst = ((long long)b * d.spatial_size + tok[k]) * d.num_heads + m;
      const int slot = atomicAdd(counts + dst, 1);
      if (kFill) ids[offsets[dst] + (unsigned)slot] = (unsigned)(s * 4 + k);
    }
  }
}

// Exclusive scan of n int32 counters into uint32 offsets (offsets[n] = total); one block.
__global__ void __launch_bounds__(1024)
msda_det_scan_kernel(const int* __restrict__ counts, unsigned* __restrict__ offsets, const long long n) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += 4096) {
    // four consecutive counters per thread
    const long long i0 = base + (long long)threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? (unsigned)counts[i0 + k] : 0u;
    const unsigned mine = v[0] + v[1] + v[2] + v[3];
    unsigned inc = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= s) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_tot[lane];
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, w, s);
        if (lane >= s) w += t;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned carry = carry_s;
    unsigned run = carry + (warp ? warp_tot[warp - 1] : 0u) + (inc - mine);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) offsets[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry_s;
}

// Ascending sort of a[0..n) by one warp: bitonic network in its "flip" form, in which every
// compare-exchange moves the smaller key to the lower index, so keys past n act as +inf and
// pairs reaching past n are simply skipped (works for any n, in shared or global memory).
__device__ __forceinline__ void warp_sort_u32(unsigned* a, const int n, const int lane) {
  for (int k = 2; (k >> 1) < n; k <<= 1) {
    for (int i = lane; i < n; i += 32) {
      const int l = i ^ (k - 1);
      if (l > i && l < n) {
        const unsigned x = a[i], y = a[l];
        if (x > y) { a[i] = y; a[l] = x; }
      }
    }
    __syncwarp();
    for (int jj = k >> 2; jj > 0; jj >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int l = i ^ jj;
        if (l > i && l < n) {
          const unsigned x = a[i], y = a[l];
          if (x > y) { a[i] = y; a[l] = x; }
        }
      }
      __syncwarp();
    }
  }
}

template <typename TV>
__device__ __forceinline__ float det_load(const TV* p) { return to_acc<float>(*p); }

// One warp per destination row (b, token, head).
template <typename TV>
__global__ void __launch_bounds__(256)
msda_det_reduce_kernel(const TV* __restrict__ grad_out, const float* __restrict__ loc,
                       const float* __restrict__ attw, const unsigned* __restrict__ offsets,
                       unsigned* __restrict__ ids, float* __restrict__ grad_value,
                       const __grid_constant__ MsdaLevels lv, const MsdaDims d, const long long n_dst) {
  __shared__ unsigned sort_buf[8][kDetSortCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int LP = d.num_levels * d.num_point;
  for (long long dst = (long long)blockIdx.x * 8 + warp; dst < n_dst; dst += (long long)gridDim.x * 8) {
    const unsigned beg = offsets[dst], end = offsets[dst + 1];
    const int n = (int)(end - beg);
    float* row = grad_value + dst * d.channels;
    if (n == 0) {
      for (int c = lane; c < d.channels; c += 32) row[c] = 0.f;
      continue;
    }
    unsigned* seg = ids + beg;
    if (n <= kDetSortCap) {
      for (int i = lane; i < n; i += 32) sort_buf[warp][i] = seg[i];
      __syncwarp();
      seg = sort_buf[warp];
    }
    warp_sort_u32(seg, n, lane);

    for (int c0 = 0; c0 < d.channels; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      for (int i0 = 0; i0 < n; i0 += 32) {
        // lane i decodes contribution i0+i: coefficient and grad_out row
        float coef = 0.f;
        long long qm = 0;
        if (i0 + lane < n) {
          const unsigned id = seg[i0 + lane];
          const long long s = id >> 2;
          const int k = id & 3;
          const int l = (int)(s % LP) / d.num_point;
          qm = s / LP;
          const float2 xy = *reinterpret_cast<const float2*>(loc + s * 2);
          int tok[4];
          float lh, lw;
          msda_sample_geom(xy.x, xy.y, lv.H[l], lv.W[l], lv.start[l], tok, lh, lw, lv.coord_fma != 0);
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float cw = k == 0 ? hh * hw : k == 1 ? hh * lw : k == 2 ? lh * hw : lh * lw;
          coef = attw[s] * cw;
        }
        const int cnt = min(32, n - i0);
        for (int i = 0; i < cnt; ++i) {
          const float cf = __shfl_sync(0xffffffffu, coef, i);
          const long long q = __shfl_sync(0xffffffffu, qm, i);
          if (c < d.channels) acc = fmaf(cf, det_load(grad_out + q * d.channels + c), acc);
        }
      }
      if (c < d.channels) row[c] = acc;
    }
    __syncwarp();
  }
}

// grad_value for fp32 / bf16 grad_out; `ws` as laid out by det_layout().
template <typename TV>
int deterministic_grad_value(cudaStream_t s, const MsdaDims& d, const MsdaLevels& lv,
                             const TV* grad_out, const float* loc, const float* attw,
                             float* grad_value, void* ws, size_t ws_bytes) {
  const DetLayout L = det_layout(d.batch, d.spatial_size, d.num_heads, d.num_levels, d.num_query, d.num_point);
  if (L.n_sample * 4 >= (1ll << 32))
    return fail(MSDA_ERR_UNSUPPORTED, "deterministic mode: %lld contributions exceed the 32-bit id space",
                L.n_sample * 4);
  if (!ws || ws_bytes < L.total)
    return fail(MSDA_ERR_WORKSPACE, "deterministic mode needs %zu workspace bytes, got %zu", L.total, ws_bytes);
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0)
    return fail(MSDA_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  char* base = static_cast<char*>(ws);
  int* counts = reinterpret_cast<int*>(base + L.counts_off);
  unsigned* offsets = reinterpret_cast<unsigned*>(base + L.offsets_off);
  unsigned* ids = reinterpret_cast<unsigned*>(base + L.ids_off);
  const size_t cnt_bytes = (size_t)(L.n_dst + 1) * 4;
  int rc = check_cuda(cudaMemsetAsync(counts, 0, cnt_bytes, s), "workspace clear");
  if (rc) return rc;
  long long blocks = (L.n_sample + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  msda_det_bin_kernel<false><<<(int)blocks, 256, 0, s>>>(loc, counts, offsets, ids, lv, d, L.n_sample);
  if ((rc = after_launch("msda_det_bin_kernel<count>"))) return rc;
  msda_det_scan_kernel<<<1, 1024, 0, s>>>(counts, offsets, L.n_dst);
  if ((rc = after_launch("msda_det_scan_kernel"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(counts, 0, cnt_bytes, s), "workspace clear"))) return rc;
  msda_det_bin_kernel<true><<<(int)blocks, 256, 0, s>>>(loc, counts, offsets, ids, lv, d, L.n_sample);
  if ((rc = after_launch("msda_det_bin_kernel<fill>"))) return rc;
  long long rblocks = (L.n_dst + 7) / 8;
  if (rblocks > 148 * 32) rblocks = 148 * 32;
  msda_det_reduce_kernel<TV><<<(int)rblocks, 256, 0, s>>>(grad_out, loc, attw, offsets, ids, grad_value,
                                                         lv, d, L.n_dst);
  return after_launch("msda_det_reduce_kernel");
}

}  // namespace msda
