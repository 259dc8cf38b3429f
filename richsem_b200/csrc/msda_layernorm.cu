// msda_layernorm.cu — residual add + LayerNorm in one pass (the encoder layer's epilogue
// around `output_proj` and around the FFN):
//
//     src = src + dropout(src2); src = norm(src)      models/richsem/deformable_transformer.py:871-872, 866-867
//
// PyTorch runs it as an add kernel (read 2, write 1 tensors of N*S*C) and a LayerNorm kernel (read 1, write 1), and
// keeps the sum alive for the backward; here the forward reads the two operands once and writes the normalised rows
// (+ 8 bytes of statistics per row), and the backward re-reads the operands instead of a saved sum:
//
//     forward   3 tensor passes (PyTorch: 5)      backward   4 tensor passes (grad_out, x, residual -> grad_in)
//
// Both are HBM-bound byte movers like the passes of msda_aux.cu: one warp per row (C = 128 * NV channels, NV float4
// per lane, a row is read with whole 512-byte warp accesses), statistics by warp shuffles, grid-stride over a grid that
// is a multiple of the SM count.  grad_gamma / grad_beta are summed per lane over the warp's rows, per block through
// shared memory, and over the blocks by a second small kernel in a fixed order (no atomics: bitwise reproducible).
// Dropout is not part of it: RichSem trains with dropout 0.0 (config/RichSem/baseline_4scale.py:42).
#include <cstdint>

#include "../../include/msda_b200.h"
#include "msda_host.h"

namespace {
using msda::after_launch;
using msda::fail;

constexpr int kLnThreads = 256;
constexpr int kLnWarps = kLnThreads / 32;
constexpr int kSms = 148;
constexpr int kLnBlocksPerSm = 6;

inline int ln_grid(long long rows) {
  long long blocks = (rows + kLnWarps - 1) / kLnWarps;
  const long long cap = (long long)kSms * kLnBlocksPerSm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// y = x + residual of one row (this lane's NV float4), all loads issued before the first use.
template <int NV>
__device__ __forceinline__ void load_sum(const float* __restrict__ x, const float* __restrict__ res, const long long row,
                                         const int lane, float4 (&y)[NV]) {
  const float4* xp = reinterpret_cast<const float4*>(x + row * (128 * NV)) + lane;
  float4 a[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = ldg_stream(xp + 32 * i);
  if (res != nullptr) {
    const float4* rp = reinterpret_cast<const float4*>(res + row * (128 * NV)) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) b[i] = ldg_stream(rp + 32 * i);
#pragma unroll
    for (int i = 0; i < NV; ++i) y[i] = add4(a[i], b[i]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) y[i] = a[i];
  }
}

// Mean and 1 / sqrt(biased variance + eps) of a row held by the warp: two passes over registers (the variance is taken
// around the mean, as torch.nn.functional.layer_norm does).
template <int NV>
__device__ __forceinline__ void row_stats(const float4 (&y)[NV], const float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (y[i].x + y[i].y) + (y[i].z + y[i].w);
  mean = warp_sum(s) * (1.f / (128 * NV));
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = y[i].x - mean, dy = y[i].y - mean, dz = y[i].z - mean, dw = y[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  rstd = rsqrtf(warp_sum(q) * (1.f / (128 * NV)) + eps);
}

template <int NV>
__global__ void __launch_bounds__(kLnThreads)
msda_add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ out, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, const long long rows, const float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kLnWarps;
  float4 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i] = gamma ? reinterpret_cast<const float4*>(gamma)[lane + 32 * i] : make_float4(1.f, 1.f, 1.f, 1.f);
    b[i] = beta ? reinterpret_cast<const float4*>(beta)[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    float4 y[NV];
    load_sum<NV>(x, res, row, lane, y);
    float mean, rstd;
    row_stats<NV>(y, eps, mean, rstd);
    float4* op = reinterpret_cast<float4*>(out + row * (128 * NV)) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      stg_stream(op + 32 * i, make_float4(fmaf((y[i].x - mean) * rstd, g[i].x, b[i].x), fmaf((y[i].y - mean) * rstd, g[i].y, b[i].y),
                                          fmaf((y[i].z - mean) * rstd, g[i].z, b[i].z), fmaf((y[i].w - mean) * rstd, g[i].w, b[i].w)));
    if (lane == 0 && mean_out != nullptr) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// grad_in = d loss / d (x + residual) (one tensor: it is the gradient of both operands);
// partial[block][0][c] = sum over the block's rows of grad_out * xhat, partial[block][1][c] = sum of grad_out.
template <int NV>
__global__ void __launch_bounds__(kLnThreads)
msda_add_layernorm_bwd_kernel(const float* __restrict__ go, const float* __restrict__ x, const float* __restrict__ res,
                              const float* __restrict__ gamma, const float* __restrict__ mean_in,
                              const float* __restrict__ rstd_in, float* __restrict__ grad_in,
                              float* __restrict__ partial, const long long rows) {
  constexpr int C = 128 * NV;
  __shared__ float4 red[kLnWarps][2 * NV][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * kLnWarps + warp;
  const long long nwarps = (long long)gridDim.x * kLnWarps;
  float4 g[NV], dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i] = gamma ? reinterpret_cast<const float4*>(gamma)[lane + 32 * i] : make_float4(1.f, 1.f, 1.f, 1.f);
    dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    float4 y[NV], d[NV];
    const float4* gp = reinterpret_cast<const float4*>(go + row * C) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) d[i] = ldg_stream(gp + 32 * i);
    const float mean = mean_in[row], rstd = rstd_in[row];
    load_sum<NV>(x, res, row, lane, y);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      // y <- xhat, d stays grad_out; gg = grad_out * gamma
      y[i] = make_float4((y[i].x - mean) * rstd, (y[i].y - mean) * rstd, (y[i].z - mean) * rstd, (y[i].w - mean) * rstd);
      const float4 gg = make_float4(d[i].x * g[i].x, d[i].y * g[i].y, d[i].z * g[i].z, d[i].w * g[i].w);
      s1 += (gg.x + gg.y) + (gg.z + gg.w);
      s2 += (gg.x * y[i].x + gg.y * y[i].y) + (gg.z * y[i].z + gg.w * y[i].w);
      dg[i] = make_float4(fmaf(d[i].x, y[i].x, dg[i].x), fmaf(d[i].y, y[i].y, dg[i].y), fmaf(d[i].z, y[i].z, dg[i].z),
                          fmaf(d[i].w, y[i].w, dg[i].w));
      db[i] = add4(db[i], d[i]);
    }
    s1 = warp_sum(s1) * (1.f / C);
    s2 = warp_sum(s2) * (1.f / C);
    float4* op = reinterpret_cast<float4*>(grad_in + row * C) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      // rstd * (gg - mean(gg) - xhat * mean(gg * xhat))
      stg_stream(op + 32 * i, make_float4(rstd * (d[i].x * g[i].x - s1 - y[i].x * s2), rstd * (d[i].y * g[i].y - s1 - y[i].y * s2),
                                          rstd * (d[i].z * g[i].z - s1 - y[i].z * s2), rstd * (d[i].w * g[i].w - s1 - y[i].w * s2)));
    }
  }
  if (partial == nullptr) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    red[warp][i][lane] = dg[i];
    red[warp][NV + i][lane] = db[i];
  }
  __syncthreads();
  // thread t sums float4 number t (and t + 256, ...) of the 2 * NV * 32 per warp, over the warps in a fixed order
  for (int k = threadIdx.x; k < 2 * NV * 32; k += kLnThreads) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) s = add4(s, red[w][k >> 5][k & 31]);
    // [which][channel]: float4 i of lane l holds channels 4 * (l + 32 * i) ...
    const int which = (k >> 5) / NV, i = (k >> 5) % NV;
    reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * 2 + which) * C)[(k & 31) + 32 * i] = s;
  }
}

// grad_gamma[c] = sum over blocks of partial[block][0][c], grad_beta[c] likewise from [1].  Block = 32 columns x 32 slices
// of the block range: every thread has its ~14 loads in flight at once, then the slices are added in index order (fixed
// order: bitwise reproducible).
__global__ void __launch_bounds__(1024)
msda_layernorm_param_grad_kernel(const float* __restrict__ partial, const int nblocks, const int C,
                                 float* __restrict__ grad_gamma, float* __restrict__ grad_beta) {
  __shared__ float part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + tx;  // (which, channel) as one index < 2 * C; 2 * C is a multiple of 32
  float s = 0.f;
#pragma unroll 4
  for (int b = ty; b < nblocks; b += 32) s += partial[(size_t)b * 2 * C + k];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) tot += part[j][tx];
    if (k < C) {
      if (grad_gamma) grad_gamma[k] = tot;
    } else if (grad_beta) {
      grad_beta[k - C] = tot;
    }
  }
}

// Blocks of the backward kernel: what is resident at once (registers bound it: 3 blocks per SM at 256 channels), so that
// there are as few partial sums as possible.
template <int NV>
int ln_bwd_grid(const long long rows) {
  static int per_sm = 0;  // benign race: every thread computes the same value
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, msda_add_layernorm_bwd_kernel<NV>, kLnThreads, 0) != cudaSuccess || n < 1) n = 2;
    per_sm = n > kLnBlocksPerSm ? kLnBlocksPerSm : n;
  }
  long long blocks = (rows + kLnWarps - 1) / kLnWarps;
  const long long cap = (long long)kSms * per_sm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_dims(const long long rows, const int channels) {
  if (rows < 0 || channels < 128 || channels > 512 || (channels & 127))
    return fail(MSDA_ERR_UNSUPPORTED, "rows=%lld channels=%d: channels must be a multiple of 128 up to 512", rows, channels);
  return MSDA_OK;
}
}  // namespace

extern "C" {

size_t msda_add_layernorm_workspace_bytes(long long rows, int channels) {
  if (rows <= 0 || channels <= 0) return 0;
  return (size_t)ln_grid(rows) * 2 * (size_t)channels * sizeof(float);
}

int msda_add_layernorm_f32(msda_stream_t stream, const float* x, const float* residual, const float* gamma,
                           const float* beta, long long rows, int channels, float eps, float* out, float* mean,
                           float* rstd) {
  if (int rc = check_dims(rows, channels)) return rc;
  if (rows == 0) return MSDA_OK;
  if (!x || !out) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if ((mean == nullptr) != (rstd == nullptr)) return fail(MSDA_ERR_INVALID_ARGUMENT, "mean and rstd go together");
  if (!aligned16(x) || !aligned16(residual) || !aligned16(gamma) || !aligned16(beta) || !aligned16(out))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
  const cudaStream_t s = (cudaStream_t)stream;
  const int grid = ln_grid(rows);
#define MSDA_LN_FWD(NV)                                                                                              \
  case NV:                                                                                                           \
    msda_add_layernorm_kernel<NV><<<grid, kLnThreads, 0, s>>>(x, residual, gamma, beta, out, mean, rstd, rows, eps); \
    break;
  switch (channels / 128) {
    MSDA_LN_FWD(1) MSDA_LN_FWD(2) MSDA_LN_FWD(3) MSDA_LN_FWD(4)
  }
#undef MSDA_LN_FWD
  return after_launch("msda_add_layernorm_kernel");
}

int msda_add_layernorm_backward_f32(msda_stream_t stream, const float* grad_out, const float* x, const float* residual,
                                    const float* gamma, const float* mean, const float* rstd, long long rows,
                                    int channels, float* grad_in, float* grad_gamma, float* grad_beta, void* workspace,
                                    size_t workspace_bytes) {
  if (int rc = check_dims(rows, channels)) return rc;
  const cudaStream_t s = (cudaStream_t)stream;
  const bool want_params = grad_gamma != nullptr || grad_beta != nullptr;
  if (rows == 0) {
    if (grad_gamma) cudaMemsetAsync(grad_gamma, 0, sizeof(float) * channels, s);
    if (grad_beta) cudaMemsetAsync(grad_beta, 0, sizeof(float) * channels, s);
    return MSDA_OK;
  }
  if (!grad_out || !x || !mean || !rstd || !grad_in) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (!aligned16(grad_out) || !aligned16(x) || !aligned16(residual) || !aligned16(gamma) || !aligned16(grad_in) ||
      !aligned16(workspace))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
  if (want_params && (!workspace || workspace_bytes < msda_add_layernorm_workspace_bytes(rows, channels)))
    return fail(MSDA_ERR_WORKSPACE, "grad_gamma / grad_beta need %zu bytes of workspace, got %zu",
                msda_add_layernorm_workspace_bytes(rows, channels), workspace_bytes);
  int grid = 1;
  float* partial = want_params ? static_cast<float*>(workspace) : nullptr;
#define MSDA_LN_BWD(NV)                                                                                     \
  case NV:                                                                                                  \
    grid = ln_bwd_grid<NV>(rows);                                                                           \
    msda_add_layernorm_bwd_kernel<NV><<<grid, kLnThreads, 0, s>>>(grad_out, x, residual, gamma, mean, rstd, \
                                                                  grad_in, partial, rows);                  \
    break;
  switch (channels / 128) {
    MSDA_LN_BWD(1) MSDA_LN_BWD(2) MSDA_LN_BWD(3) MSDA_LN_BWD(4)
  }
#undef MSDA_LN_BWD
  if (int rc = after_launch("msda_add_layernorm_bwd_kernel")) return rc;
  if (want_params) {
    msda_layernorm_param_grad_kernel<<<2 * channels / 32, 1024, 0, s>>>(partial, grid, channels, grad_gamma, grad_beta);
    return after_launch("msda_layernorm_param_grad_kernel");
  }
  return MSDA_OK;
}

}  // extern "C"
