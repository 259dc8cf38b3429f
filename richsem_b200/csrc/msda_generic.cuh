// msda_generic.cuh — shape-agnostic kernels: any channels, any L, P; fp32 or fp64.
//
// These serve everything the tuned D=32 kernels do not (the reference's test sweep uses
// channels 30, 64, 71, 1025, 2048, 3096 and fp64 gradcheck, models/richsem/ops/test.py:63-86).
// One warp owns one (batch, query, head); lanes stride over channels, so no
// channel-count-specific variants are needed (the reference carries seven backward
// variants, cuh:301-920).  grad_sampling_loc / grad_attn_weight are warp-shuffle
// reduced and written once; grad_value uses one atomicAdd per element.
#pragma once

#include <cuda_bf16.h>

#include "msda_common.cuh"

namespace msda {

template <typename T, typename TV>
__device__ __forceinline__ T to_acc(TV v) { return (T)v; }
template <>
__device__ __forceinline__ float to_acc<float, __nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename TV, typename T>
__device__ __forceinline__ TV from_acc(T v) { return (TV)v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// TV: storage type of value / out / grad_out (float, double, __nv_bfloat16)
// T : arithmetic type, and storage type of sampling_loc / attn_weight / all gradients
template <typename TV, typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const TV* __restrict__ value, const T* __restrict__ loc,
                        const T* __restrict__ attw, TV* __restrict__ out,
                        const __grid_constant__ MsdaLevels lv, const MsdaDims d) {
  const int lane = threadIdx.x & 31;
  const long long n_task = (long long)d.batch * d.num_query * d.num_heads;
  const int LP = d.num_levels * d.num_point;
  for (long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); task < n_task;
       task += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int m = (int)(task % d.num_heads);
    const int b = (int)(task / ((long long)d.num_heads * d.num_query));
    const TV* value_b = value + (size_t)b * d.spatial_size * d.num_heads * d.channels;
    const T* loc_t = loc + (size_t)task * LP * 2;
    const T* w_t = attw + (size_t)task * LP;
    for (int c0 = 0; c0 < d.channels; c0 += 32) {
      const int c = c0 + lane;
      T acc = T(0);
      for (int l = 0; l < d.num_levels; ++l) {
        const int H = lv.H[l], W = lv.W[l], st = lv.start[l];
        for (int p = 0; p < d.num_point; ++p) {
          const int lp = l * d.num_point + p;
          const T x = loc_t[lp * 2], y = loc_t[lp * 2 + 1], a = w_t[lp];
          int tok[4];
          T lh, lw;
          if (!msda_sample_geom(x, y, H, W, st, tok, lh, lw, lv.coord_fma != 0)) continue;
          if (c < d.channels) {
            const T hh = T(1) - lh, hw = T(1) - lw;
            const size_t ch = (size_t)m * d.channels + c;
            const size_t rs = (size_t)d.num_heads * d.channels;
            const T v0 = tok[0] >= 0 ? to_acc<T>(value_b[tok[0] * rs + ch]) : T(0);
            const T v1 = tok[1] >= 0 ? to_acc<T>(value_b[tok[1] * rs + ch]) : T(0);
            const T v2 = tok[2] >= 0 ? to_acc<T>(value_b[tok[2] * rs + ch]) : T(0);
            const T v3 = tok[3] >= 0 ? to_acc<T>(value_b[tok[3] * rs + ch]) : T(0);
            acc += a * ((hh * hw) * v0 + (hh * lw) * v1 + (lh * hw) * v2 + (lh * lw) * v3);
          }
        }
      }
      if (c < d.channels) out[(size_t)task * d.channels + c] = from_acc<TV>(acc);
    }
  }
}

// kScatter=false skips the grad_value atomics (deterministic mode computes it separately).
template <typename TV, typename T, bool kScatter>
__global__ void __launch_bounds__(256)
msda_bwd_generic_kernel(const TV* __restrict__ grad_out, const TV* __restrict__ value,
                        const T* __restrict__ loc, const T* __restrict__ attw,
                        T* __restrict__ grad_value, T* __restrict__ grad_loc,
                        T* __restrict__ grad_attw, const __grid_constant__ MsdaLevels lv,
                        const MsdaDims d) {
  const int lane = threadIdx.x & 31;
  const long long n_task = (long long)d.batch * d.num_query * d.num_heads;
  const int LP = d.num_levels * d.num_point;
  const size_t rs = (size_t)d.num_heads * d.channels;
  for (long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); task < n_task;
       task += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int m = (int)(task % d.num_heads);
    const int b = (int)(task / ((long long)d.num_heads * d.num_query));
    const TV* value_b = value + (size_t)b * d.spatial_size * rs;
    T* gvalue_b = grad_value + (size_t)b * d.spatial_size * rs;
    const T* loc_t = loc + (size_t)task * LP * 2;
    const T* w_t = attw + (size_t)task * LP;
    const TV* go_t = grad_out + (size_t)task * d.channels;
    for (int l = 0; l < d.num_levels; ++l) {
      const int H = lv.H[l], W = lv.W[l], st = lv.start[l];
      for (int p = 0; p < d.num_point; ++p) {
        const int lp = l * d.num_point + p;
        const T x = loc_t[lp * 2], y = loc_t[lp * 2 + 1], a = w_t[lp];
        int tok[4];
        T lh, lw;
        T ga = T(0), gx = T(0), gy = T(0);
        if (msda_sample_geom(x, y, H, W, st, tok, lh, lw, lv.coord_fma != 0)) {  // warp-uniform
          const T hh = T(1) - lh, hw = T(1) - lw;
          for (int c = lane; c < d.channels; c += 32) {
            const size_t ch = (size_t)m * d.channels + c;
            const T g = to_acc<T>(go_t[c]);
            const T t = g * a;
            T v0 = T(0), v1 = T(0), v2 = T(0), v3 = T(0);
            if (tok[0] >= 0) { v0 = to_acc<T>(value_b[tok[0] * rs + ch]); if (kScatter) atomicAdd(gvalue_b + tok[0] * rs + ch, (hh * hw) * t); }
            if (tok[1] >= 0) { v1 = to_acc<T>(value_b[tok[1] * rs + ch]); if (kScatter) atomicAdd(gvalue_b + tok[1] * rs + ch, (hh * lw) * t); }
            if (tok[2] >= 0) { v2 = to_acc<T>(value_b[tok[2] * rs + ch]); if (kScatter) atomicAdd(gvalue_b + tok[2] * rs + ch, (lh * hw) * t); }
            if (tok[3] >= 0) { v3 = to_acc<T>(value_b[tok[3] * rs + ch]); if (kScatter) atomicAdd(gvalue_b + tok[3] * rs + ch, (lh * lw) * t); }
            ga += g * ((hh * hw) * v0 + (hh * lw) * v1 + (lh * hw) * v2 + (lh * lw) * v3);
            gx += t * (hh * (v1 - v0) + lh * (v3 - v2));
            gy += t * (hw * (v2 - v0) + lw * (v3 - v1));
          }
          ga = warp_sum(ga);
          gx = warp_sum(gx) * T(W);
          gy = warp_sum(gy) * T(H);
        }
        if (lane == 0) {
          grad_attw[(size_t)task * LP + lp] = ga;
          grad_loc[((size_t)task * LP + lp) * 2] = gx;
          grad_loc[((size_t)task * LP + lp) * 2 + 1] = gy;
        }
      }
    }
  }
}

// Index-contract probe: corner token indices for every sample (see msda_b200.h).
__global__ void __launch_bounds__(256)
msda_corners_kernel(const float* __restrict__ loc, int* __restrict__ corners,
                    const __grid_constant__ MsdaLevels lv, const long long n_sample,
                    const int num_levels, const int num_point) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_sample;
       i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((i / num_point) % num_levels);
    int tok[4];
    float lh, lw;
    msda_sample_geom(loc[i * 2], loc[i * 2 + 1], lv.H[l], lv.W[l], lv.start[l], tok, lh, lw, lv.coord_fma != 0);
    reinterpret_cast<int4*>(corners)[i] = make_int4(tok[0], tok[1], tok[2], tok[3]);
  }
}

}  // namespace msda
