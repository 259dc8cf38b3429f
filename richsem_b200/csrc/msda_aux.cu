// msda_aux.cu — the elementwise passes either side of the sampling core:
//
//   * value preparation (models/richsem/ops/modules/ms_deform_attn.py:94-97): padding-mask zeroing of the
//     projected value, fused with the bf16 cast the bf16 kernels want — one pass instead of masked_fill + to();
//     for fp32 the rows are zeroed IN PLACE, so only the masked rows are written at all.
//   * two-stage proposals (models/richsem/utils.py:10-65, gen_encoder_output_proposals): per-token anchor
//     boxes in logit space + masked copy of the encoder memory — one pass over S tokens instead of ~25
//     PyTorch kernels per call.
//
// All kernels are HBM-bound byte movers: one warp per token row, 16-byte streaming accesses, grid-stride
// over a grid that is a multiple of the SM count.
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "msda_host.h"

namespace {
using msda::after_launch;
using msda::check_cuda;
using msda::fail;

constexpr int kAuxThreads = 256;
constexpr int kAuxBlocksPerSm = 8;
constexpr int kSms = 148;

inline int aux_grid(long long warps_needed) {
  long long blocks = (warps_needed + kAuxThreads / 32 - 1) / (kAuxThreads / 32);
  const long long cap = (long long)kSms * kAuxBlocksPerSm;
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
__device__ __forceinline__ void stg_stream(uint2* p, uint2 v) {
  asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // round-to-nearest-even, as Tensor.to(bfloat16)
  return r;
}

// ---- value preparation ------------------------------------------------------------------------------
// out[row, :] = mask[row] ? 0 : bf16(in[row, :]).  One warp per row; row_elems % 4 == 0.  Masked rows are
// not read.
__global__ void __launch_bounds__(kAuxThreads)
msda_value_prepare_bf16_kernel(const float* __restrict__ in, const uint8_t* __restrict__ mask,
                               uint16_t* __restrict__ out, long long rows, int row_elems) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kAuxThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kAuxThreads / 32);
  const int quads = row_elems >> 2;
  // two rows per warp and iteration, every load issued before the first store: 4 x 16 bytes in flight per lane at
  // 256 channels
  for (long long row = warp0; row < rows; row += 2 * nwarps) {
    const long long row2 = row + nwarps;
    const bool has2 = row2 < rows;
    const bool live1 = !(mask != nullptr && mask[row] != 0);
    const bool live2 = has2 && !(mask != nullptr && mask[row2] != 0);
    const float4* src1 = reinterpret_cast<const float4*>(in + row * row_elems);
    const float4* src2 = reinterpret_cast<const float4*>(in + row2 * row_elems);
    uint2* dst1 = reinterpret_cast<uint2*>(out + row * row_elems);
    uint2* dst2 = reinterpret_cast<uint2*>(out + row2 * row_elems);
    for (int i = lane; i < quads; i += 64) {
      const bool b = i + 32 < quads;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v1a = live1 ? ldg_stream(src1 + i) : z;
      const float4 v1b = (live1 && b) ? ldg_stream(src1 + i + 32) : z;
      const float4 v2a = live2 ? ldg_stream(src2 + i) : z;
      const float4 v2b = (live2 && b) ? ldg_stream(src2 + i + 32) : z;
      stg_stream(dst1 + i, make_uint2(pack_bf16x2(v1a.x, v1a.y), pack_bf16x2(v1a.z, v1a.w)));
      if (b) stg_stream(dst1 + i + 32, make_uint2(pack_bf16x2(v1b.x, v1b.y), pack_bf16x2(v1b.z, v1b.w)));
      if (has2) {
        stg_stream(dst2 + i, make_uint2(pack_bf16x2(v2a.x, v2a.y), pack_bf16x2(v2a.z, v2a.w)));
        if (b) stg_stream(dst2 + i + 32, make_uint2(pack_bf16x2(v2b.x, v2b.y), pack_bf16x2(v2b.z, v2b.w)));
      }
    }
  }
}

// data[row, :] = 0 where mask[row]; only masked rows are touched.  Padding is a rectangle per level, i.e. long runs
// of masked rows, so rows are dealt to warps round-robin (a warp that owned 32 consecutive rows would do all the
// work of its run alone: 9 % occupancy in the first version); a warp tests four of its rows per iteration.
__global__ void __launch_bounds__(kAuxThreads)
msda_zero_masked_rows_kernel(float* __restrict__ data, const uint8_t* __restrict__ mask, long long rows,
                             int row_elems) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kAuxThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kAuxThreads / 32);
  const int quads = row_elems >> 2;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = warp0; row < rows; row += 4 * nwarps) {
    bool flag[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = row + u * nwarps;
      flag[u] = r < rows && mask[r] != 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (flag[u]) {
        float4* dst = reinterpret_cast<float4*>(data + (row + u * nwarps) * row_elems);
        for (int i = lane; i < quads; i += 32) dst[i] = z;
      }
    }
  }
}

// dst[row] = keep ? src[row] : 0 for one row of `quads` float4; the loads of a 64-quad chunk are issued before its stores.
__device__ __forceinline__ void copy_row_or_zero(const float4* __restrict__ src, float4* __restrict__ dst, int quads,
                                                 int lane, bool keep) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane; i < quads; i += 128) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (keep && i + 32 * u < quads) ? ldg_stream(src + i + 32 * u) : z;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + 32 * u < quads) stg_stream(dst + i + 32 * u, v[u]);
  }
}

// ---- two-stage proposals ----------------------------------------------------------------------------
struct AuxLevels {
  int H[MSDA_MAX_LEVELS];
  int W[MSDA_MAX_LEVELS];
  int start[MSDA_MAX_LEVELS];  // running sum of H*W (utils.py:25,54 `_cur`)
  int num_levels;
};

// valid_hw[(b*L + l)*2 + {0,1}] = (#unmasked tokens in the first column, #unmasked tokens in the first row)
// of level l of image b (utils.py:27-28).  One warp per (image, level).
__device__ __forceinline__ int2 valid_hw_of(const uint8_t* __restrict__ mask, const AuxLevels& lv, int task,
                                            int spatial_size, int lane) {
  const int b = task / lv.num_levels, l = task - b * lv.num_levels;
  const int H = lv.H[l], W = lv.W[l];
  const uint8_t* m = mask + (long long)b * spatial_size + lv.start[l];
  int vh = 0, vw = 0;
  for (int h = lane; h < H; h += 32) vh += m[(long long)h * W] == 0;
  for (int w = lane; w < W; w += 32) vw += m[w] == 0;
  for (int o = 16; o; o >>= 1) {
    vh += __shfl_xor_sync(0xffffffffu, vh, o);
    vw += __shfl_xor_sync(0xffffffffu, vw, o);
  }
  return make_int2(vh, vw);
}

// Stand-alone pass for batches too large for the in-block table of msda_proposals_kernel.
__global__ void msda_valid_hw_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ valid_hw,
                                     const __grid_constant__ AuxLevels lv, int batch, int spatial_size) {
  const int lane = threadIdx.x & 31;
  const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= batch * lv.num_levels) return;
  const int2 v = valid_hw_of(mask, lv, task, spatial_size, lane);
  if (lane == 0) {
    valid_hw[2 * task] = v.x;
    valid_hw[2 * task + 1] = v.y;
  }
}

constexpr int kValidInBlock = 128;  // (image, level) pairs a block recomputes for itself (L2-resident mask bytes)

// One warp per token: proposal (cx, cy, w, h) of utils.py:30-46 in logit space (:50-52), validity test (:49),
// masked copy of the memory row (:54-56).  Same fp32 operations in the same order as the PyTorch expressions.
__global__ void __launch_bounds__(kAuxThreads)
msda_proposals_kernel(const float* __restrict__ memory, const uint8_t* __restrict__ mask,
                      const int32_t* __restrict__ valid_hw, const float* __restrict__ wh_base,
                      float* __restrict__ out_memory, float* __restrict__ out_proposals,
                      const __grid_constant__ AuxLevels lv, int batch, int spatial_size, int channels) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kAuxThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kAuxThreads / 32);
  const long long rows = (long long)batch * spatial_size;
  const int quads = channels >> 2;
  const float base_w = wh_base ? wh_base[0] : 0.05f;
  const float base_h = wh_base ? wh_base[1] : 0.05f;
  const float inf = __int_as_float(0x7f800000);
  // valid H / W of every (image, level): small batches recompute the table per block (a few hundred mask bytes out of
  // L2) instead of waiting for a separate launch; large ones read the table msda_valid_hw_kernel left in valid_hw
  __shared__ int s_valid[2 * kValidInBlock];
  const int pairs = batch * lv.num_levels;
  const bool in_block = mask != nullptr && pairs <= kValidInBlock;
  if (in_block) {
    for (int task = threadIdx.x >> 5; task < pairs; task += kAuxThreads / 32) {
      const int2 v = valid_hw_of(mask, lv, task, spatial_size, lane);
      if (lane == 0) {
        s_valid[2 * task] = v.x;
        s_valid[2 * task + 1] = v.y;
      }
    }
    __syncthreads();
  }
  // The per-token arithmetic (level lookup, integer division, IEEE divisions, logf) costs ~250 instructions, and with
  // one token per warp the kernel was bound by instruction issue (69 % of issue slots, 13.4 M warp instructions:
  // profiles/r1_aux_passes.md).  So a warp takes FOUR tokens at a time, 8 lanes each: the arithmetic is SIMT-shared
  // between the four, and a lane moves 8 x 16 bytes of its row (all loads ahead of the stores) at 256 channels.
  constexpr int kLanesPerRow = 8, kRowsPerWarp = 32 / kLanesPerRow;
  const int sub = lane / kLanesPerRow, j = lane % kLanesPerRow;
  const long long first = warp0 * kRowsPerWarp + sub;
  const long long stride = nwarps * kRowsPerWarp;
  // (image, token) of this lane group's row, advanced incrementally (no 64-bit division per row)
  int b = (int)(first / spatial_size);
  int s = (int)(first - (long long)b * spatial_size);
  const int step_b = (int)(stride / spatial_size), step_s = (int)(stride - (long long)step_b * spatial_size);
  for (long long row0 = warp0 * kRowsPerWarp; row0 < rows; row0 += stride) {
    const long long row = row0 + sub;
    const bool live = row < rows;
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < lv.num_levels; ++k) l += s >= lv.start[k];
    const int W = lv.W[l];
    const int rel = s - lv.start[l];
    const int y = (int)((unsigned)rel / (unsigned)W), x = rel - y * W;
    float vw = (float)W, vh = (float)lv.H[l];
    bool masked = false;
    if (mask != nullptr && live) {
      masked = mask[row] != 0;
      const int pair = b * lv.num_levels + l;
      vh = (float)(in_block ? s_valid[2 * pair] : valid_hw[2 * pair]);
      vw = (float)(in_block ? s_valid[2 * pair + 1] : valid_hw[2 * pair + 1]);
    }
    const float scale_l = (float)(1 << l);  // 2.0 ** lvl (utils.py:41,43)
    // lanes 0..3 of a group take one component each: cx, cy, w, h
    const int comp = j & 3;
    const float centre = __fdiv_rn((float)(comp == 0 ? x : y) + 0.5f, comp == 0 ? vw : vh);  // one division per lane
    const float extent = __fmul_rn(comp == 2 ? base_w : base_h, scale_l);
    const float pk = comp < 2 ? centre : extent;
    const bool ok = (pk > 0.01f) && (pk < 0.99f);
    const bool valid = ((__ballot_sync(0xffffffffu, ok) >> (sub * kLanesPerRow)) & 0xfu) == 0xfu;
    const bool keep = valid && !masked;
    if (live) {
      if (j < 4) out_proposals[4 * row + j] = keep ? logf(__fdiv_rn(pk, __fsub_rn(1.f, pk))) : inf;
      const float4* src = reinterpret_cast<const float4*>(memory + row * channels);
      float4* dst = reinterpret_cast<float4*>(out_memory + row * channels);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = j; i < quads; i += 8 * kLanesPerRow) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = (keep && i + kLanesPerRow * u < quads) ? ldg_stream(src + i + kLanesPerRow * u) : z;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i + kLanesPerRow * u < quads) stg_stream(dst + i + kLanesPerRow * u, v[u]);
      }
    }
    b += step_b;
    s += step_s;
    if (s >= spatial_size) {
      s -= spatial_size;
      ++b;
    }
  }
}

// grad_memory[row, :] = kept(row) ? grad_output_memory[row, :] : 0, kept(row) <=> the row's proposal is finite.
__global__ void __launch_bounds__(kAuxThreads)
msda_proposals_backward_kernel(const float* __restrict__ grad_out, const float* __restrict__ proposals,
                               float* __restrict__ grad_memory, long long rows, int channels) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kAuxThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kAuxThreads / 32);
  const int quads = channels >> 2;
  for (long long row = warp0; row < rows; row += nwarps) {
    const bool keep = isfinite(proposals[4 * row]);
    copy_row_or_zero(reinterpret_cast<const float4*>(grad_out + row * channels),
                     reinterpret_cast<float4*>(grad_memory + row * channels), quads, lane, keep);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int resolve_aux_levels(cudaStream_t stream, const int64_t* shapes_dev, int num_levels, int spatial_size,
                       const msda_opts* opts, AuxLevels* lv) {
  if (num_levels < 1 || num_levels > MSDA_MAX_LEVELS)
    return fail(MSDA_ERR_UNSUPPORTED, "num_levels=%d outside [1,%d]", num_levels, MSDA_MAX_LEVELS);
  if (opts && opts->struct_size != sizeof(msda_opts))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_opts.struct_size=%u, library expects %zu", opts->struct_size,
                sizeof(msda_opts));
  int64_t shp[2 * MSDA_MAX_LEVELS];
  if (opts && opts->spatial_shapes_host) {
    memcpy(shp, opts->spatial_shapes_host, sizeof(int64_t) * 2 * num_levels);
  } else {
    if (!shapes_dev) return fail(MSDA_ERR_INVALID_ARGUMENT, "spatial_shapes is NULL");
    cudaError_t e =
        cudaMemcpyAsync(shp, shapes_dev, sizeof(int64_t) * 2 * num_levels, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return check_cuda(e, "copy of spatial_shapes to the host");
  }
  memset(lv, 0, sizeof(*lv));
  lv->num_levels = num_levels;
  long long cur = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int64_t H = shp[2 * l], W = shp[2 * l + 1];
    if (H < 1 || W < 1 || H * W > INT32_MAX)
      return fail(MSDA_ERR_INVALID_ARGUMENT, "level %d: bad shape (%lld,%lld)", l, (long long)H, (long long)W);
    lv->H[l] = (int)H;
    lv->W[l] = (int)W;
    lv->start[l] = (int)cur;
    cur += H * W;
  }
  if (cur != spatial_size)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "sum of H*W over the levels is %lld, spatial_size is %d", cur, spatial_size);
  return MSDA_OK;
}
}  // namespace

extern "C" {

int msda_value_prepare_bf16(msda_stream_t stream, const float* projected, const uint8_t* padding_mask,
                            long long rows, int row_elems, uint16_t* value_out) {
  if (rows < 0 || row_elems < 4 || (row_elems & 3))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "rows=%lld row_elems=%d (row_elems must be a positive multiple of 4)", rows,
                row_elems);
  if (rows == 0) return MSDA_OK;
  if (!projected || !value_out) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (!aligned16(projected) || (reinterpret_cast<uintptr_t>(value_out) & 7u))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "projected must be 16-byte aligned and value_out 8-byte aligned");
  msda_value_prepare_bf16_kernel<<<aux_grid(rows), kAuxThreads, 0, (cudaStream_t)stream>>>(
      projected, padding_mask, value_out, rows, row_elems);
  return after_launch("msda_value_prepare_bf16_kernel");
}

int msda_zero_masked_rows_f32(msda_stream_t stream, float* data, const uint8_t* padding_mask, long long rows,
                              int row_elems) {
  if (rows < 0 || row_elems < 4 || (row_elems & 3))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "rows=%lld row_elems=%d (row_elems must be a positive multiple of 4)", rows,
                row_elems);
  if (rows == 0) return MSDA_OK;
  if (!data || !padding_mask) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (!aligned16(data)) return fail(MSDA_ERR_INVALID_ARGUMENT, "data must be 16-byte aligned");
  msda_zero_masked_rows_kernel<<<aux_grid((rows + 3) / 4), kAuxThreads, 0, (cudaStream_t)stream>>>(
      data, padding_mask, rows, row_elems);
  return after_launch("msda_zero_masked_rows_kernel");
}

int msda_encoder_proposals_f32(msda_stream_t stream, const float* memory, const uint8_t* padding_mask,
                               const int64_t* spatial_shapes, const float* wh_base, int batch, int spatial_size,
                               int channels, int num_levels, float* output_memory, float* output_proposals,
                               int32_t* valid_hw_workspace, const msda_opts* opts) {
  cudaStream_t s = (cudaStream_t)stream;
  if (batch < 0 || spatial_size < 1 || channels < 4 || (channels & 3))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "bad dimensions: batch=%d spatial_size=%d channels=%d (multiple of 4)",
                batch, spatial_size, channels);
  AuxLevels lv;
  const int rc = resolve_aux_levels(s, spatial_shapes, num_levels, spatial_size, opts, &lv);
  if (rc != MSDA_OK) return rc;
  if (batch == 0) return MSDA_OK;
  if (!memory || !output_memory || !output_proposals) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (!aligned16(memory) || !aligned16(output_memory) || !aligned16(output_proposals))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "memory / output_memory / output_proposals must be 16-byte aligned");
  if (padding_mask && batch * num_levels > kValidInBlock) {
    if (!valid_hw_workspace)
      return fail(MSDA_ERR_WORKSPACE, "a padding mask needs valid_hw_workspace (batch * num_levels * 2 int32)");
    const int tasks = batch * num_levels;
    msda_valid_hw_kernel<<<(tasks + 3) / 4, 128, 0, s>>>(padding_mask, valid_hw_workspace, lv, batch, spatial_size);
    const int rc2 = after_launch("msda_valid_hw_kernel");
    if (rc2 != MSDA_OK) return rc2;
  }
  msda_proposals_kernel<<<aux_grid(((long long)batch * spatial_size + 3) / 4), kAuxThreads, 0, s>>>(
      memory, padding_mask, valid_hw_workspace, wh_base, output_memory, output_proposals, lv, batch, spatial_size,
      channels);
  return after_launch("msda_proposals_kernel");
}

int msda_encoder_proposals_backward_f32(msda_stream_t stream, const float* grad_output_memory,
                                        const float* output_proposals, int batch, int spatial_size, int channels,
                                        float* grad_memory) {
  if (batch < 0 || spatial_size < 1 || channels < 4 || (channels & 3))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "bad dimensions: batch=%d spatial_size=%d channels=%d (multiple of 4)",
                batch, spatial_size, channels);
  if (batch == 0) return MSDA_OK;
  if (!grad_output_memory || !output_proposals || !grad_memory)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (!aligned16(grad_output_memory) || !aligned16(grad_memory) || !aligned16(output_proposals))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
  const long long rows = (long long)batch * spatial_size;
  msda_proposals_backward_kernel<<<aux_grid(rows), kAuxThreads, 0, (cudaStream_t)stream>>>(
      grad_output_memory, output_proposals, grad_memory, rows, channels);
  return after_launch("msda_proposals_backward_kernel");
}

}  // extern "C"

// =====================================================================================================
// Two-stage query selection (second half): models/richsem/deformable_transformer.py:367-369
//   topk_proposals = torch.topk(enc_outputs_class_unselected.max(-1)[0], num_queries, dim=1)[1]
// msda_rowmax_f32: scores[row] = max over the class logits of a token (HBM-bound: reads rows x K floats once).
// msda_topk_rows_f32: per image, the indices of the k largest scores, sorted by descending score (ties: lower
// index first; NaN counts as the largest value, like torch.topk) — one thread block per image: 4-pass 8-bit radix
// select of the k-th key over the L2-resident scores, ordered collection, bitonic sort of the k winners.
// =====================================================================================================
namespace {

__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// torch.max semantics: a NaN anywhere in the row makes the result NaN.
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }

// kLanes lanes per row (32 / kLanes rows per warp): 32 for wide rows (LVIS: 1203 classes), 8 for narrow ones (COCO: 91),
// where a whole warp per row would leave most lanes without a 16-byte load.
template <int kLanes>
__global__ void __launch_bounds__(kAuxThreads)
msda_rowmax_kernel(const float* __restrict__ logits, float* __restrict__ scores, long long rows, int K) {
  constexpr int kRowsPerWarp = 32 / kLanes;
  const int lane = threadIdx.x & 31, sub = lane / kLanes, j = lane % kLanes;
  const long long warp0 = (long long)blockIdx.x * (kAuxThreads / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (kAuxThreads / 32);
  for (long long row0 = warp0 * kRowsPerWarp; row0 < rows; row0 += nwarps * kRowsPerWarp) {
    const long long row = row0 + sub;
    float m = -INFINITY;
    if (row < rows) {
      // K is arbitrary (91, 1203, ...), so rows start at any 4-byte offset: up to 3 scalars in front, then the
      // 16-byte aligned body with four 16-byte loads in flight per lane, then up to 3 scalars behind
      const float* src = logits + row * K;
      int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) >> 2);
      head = head < K ? head : K;
      const int nq = (K - head) >> 2;
      const int tail0 = head + 4 * nq;
      if (j < head) m = ldg_stream_f1(src + j);
      if (j < 3 && tail0 + j < K) m = max_nan(m, ldg_stream_f1(src + tail0 + j));
      const float4* body = reinterpret_cast<const float4*>(src + head);
      for (int i = j; i < nq; i += 4 * kLanes) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          v[u] = (i + kLanes * u < nq) ? ldg_stream(body + i + kLanes * u)
                                       : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int u = 0; u < 4; ++u) m = max_nan(max_nan(m, max_nan(v[u].x, v[u].y)), max_nan(v[u].z, v[u].w));
      }
    }
#pragma unroll
    for (int o = kLanes / 2; o; o >>= 1) m = max_nan(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (j == 0 && row < rows) scores[row] = m;
  }
}

// Monotone map float -> uint32 (larger float = larger key); every NaN maps to the largest key.
__device__ __forceinline__ unsigned topk_key(float f) {
  if (f != f) return 0xffffffffu;
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kTopkThreads = 1024;
constexpr int kTopkMax = 1024;
constexpr int kTopkStageMax = 49152;  // keys of one row staged in shared memory (192 KB of the 227 KB)

__global__ void __launch_bounds__(kTopkThreads)
msda_topk_rows_kernel(const float* __restrict__ scores, long long* __restrict__ indices, float* __restrict__ values,
                      int row_len, int k, bool staged) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_want, s_count_gt, s_tie_base;
  __shared__ unsigned s_warp_sum[kTopkThreads / 32];
  __shared__ unsigned long long s_list[kTopkMax];
  extern __shared__ unsigned s_keys[];  // staged == true: the row's keys (the five passes below then never leave the SM)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = scores + (long long)blockIdx.x * row_len;
  if (staged) {
    for (int i = tid; i < row_len; i += 4 * kTopkThreads) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * kTopkThreads < row_len) ? row[i + u * kTopkThreads] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * kTopkThreads < row_len) s_keys[i + u * kTopkThreads] = topk_key(v[u]);
    }
    __syncthreads();
  }

  // ---- radix select: key of the k-th largest element ------------------------------------------------
  if (tid == 0) {
    s_prefix = 0u;
    s_want = (unsigned)k;
  }
  unsigned prefix_mask = 0u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    for (int i0 = 0; i0 < row_len; i0 += kTopkThreads) {  // uniform trip count: whole warps reach the match
      const int i = i0 + tid;
      unsigned bin = 256u;  // "not a candidate"
      if (i < row_len) {
        const unsigned key = staged ? s_keys[i] : topk_key(row[i]);
        if ((key & prefix_mask) == prefix) bin = (key >> shift) & 255u;
      }
      // scores cluster in a few exponent bins, so the lanes of a warp mostly hit the same counter: one atomic per
      // distinct bin and warp
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin < 256u && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns bins 8l .. 8l+7; suffix sums from the top bin down
      unsigned h[8], mine = 0u;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        h[b] = hist[8 * lane + b];
        mine += h[b];
      }
      // inclusive suffix sum over the lanes: elements in this lane's bins and in all higher bins
      unsigned suffix = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += t;
      }
      const unsigned above = suffix - mine;
      const unsigned want = s_want;
      __syncwarp();  // every lane has read s_want before the owning lane rewrites it
      if (above < want && want <= suffix) {  // the k-th element falls into one of this lane's bins (exactly one lane)
        unsigned cum = above, cum_at = 0u;
        int chosen = -1;
#pragma unroll
        for (int b = 7; b >= 0; --b) {
          if (chosen < 0) {
            if (want <= cum + h[b]) {
              chosen = b;
              cum_at = cum;
            } else {
              cum += h[b];
            }
          }
        }
        s_prefix = prefix | ((unsigned)(8 * lane + chosen) << shift);
        s_want = want - cum_at;
      }
    }
    prefix_mask |= 255u << shift;
    __syncthreads();
  }
  const unsigned T = s_prefix;      // key of the k-th largest element
  const unsigned ties = s_want;     // how many elements equal to T belong to the top k
  const unsigned n_gt = (unsigned)k - ties;

  // ---- collection: everything above T, and the first `ties` elements equal to T in index order -------
  if (tid == 0) {
    s_count_gt = 0u;
    s_tie_base = 0u;
  }
  for (int e = tid; e < kTopkMax; e += kTopkThreads) s_list[e] = 0ull;
  __syncthreads();
  for (int i0 = 0; i0 < row_len; i0 += kTopkThreads) {
    const int i = i0 + tid;
    unsigned key = 0u;
    bool gt = false, eq = false;
    if (i < row_len) {
      key = staged ? s_keys[i] : topk_key(row[i]);
      gt = key > T;
      eq = key == T;
    }
    if (gt) {
      const unsigned slot = atomicAdd(&s_count_gt, 1u);
      s_list[slot] = ((unsigned long long)key << 32) | (0xffffffffu - (unsigned)i);
    }
    // ordered rank among the elements equal to T (rare: one barrier tells whether this chunk has any)
    const int chunk_ties = __syncthreads_count(eq);
    if (chunk_ties == 0 || s_tie_base >= ties) continue;   // uniform: both operands are block-wide values
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    const unsigned before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp_sum[warp] = __popc(bal);
    __syncthreads();
    unsigned base = s_tie_base;
    for (int w = 0; w < warp; ++w) base += s_warp_sum[w];
    if (eq) {
      const unsigned r = base + before;
      if (r < ties) s_list[n_gt + r] = ((unsigned long long)key << 32) | (0xffffffffu - (unsigned)i);
    }
    __syncthreads();
    if (tid == 0) s_tie_base += (unsigned)chunk_ties;
    __syncthreads();
  }
  __syncthreads();

  // ---- bitonic sort, descending by (key, -index) ------------------------------------------------------
  // (the element lives in a register; partners less than a warp apart are exchanged with shuffles, so only 15 of the
  // 55 compare-exchange stages need shared memory and barriers)
  unsigned long long v = s_list[tid];
  for (int size = 2; size <= kTopkMax; size <<= 1) {
    const bool desc = (tid & size) == 0;
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      unsigned long long other;
      if (stride >= 32) {
        s_list[tid] = v;
        __syncthreads();
        other = s_list[tid ^ stride];
        __syncthreads();
      } else {
        other = __shfl_xor_sync(0xffffffffu, v, stride);
      }
      const bool lower = (tid & stride) == 0;
      const bool keep_max = lower == desc;
      v = keep_max ? (v > other ? v : other) : (v < other ? v : other);
    }
  }
  s_list[tid] = v;
  __syncthreads();
  if (tid < k) {
    const unsigned long long e = s_list[tid];
    const unsigned idx = 0xffffffffu - (unsigned)(e & 0xffffffffull);
    indices[(long long)blockIdx.x * k + tid] = (long long)idx;
    if (values != nullptr) values[(long long)blockIdx.x * k + tid] = row[idx];
  }
}

// ---- the same selection by a thread-block CLUSTER per image ------------------------------------------------
// One block per image is latency / issue-bound on a single SM (37-41 us for 22,223 scores: 165 k warp instructions
// on one SM, profiles/r1_aux_passes.md).  Here kTopkCluster blocks (one cluster, 8 SMs) share an image: each stages
// 1/8 of the keys in its own shared memory and histograms them; the per-block histograms are merged through
// distributed shared memory (every block reads the eight 256-bin tables and takes the same decision), winners are
// appended straight into block 0's list over DSMEM, and block 0 sorts them.
constexpr int kTopkCluster = 8;
constexpr int kTopkSliceMax = 12288;  // keys per block staged in shared memory (48 KB)

__global__ void __launch_bounds__(kTopkThreads)
msda_topk_rows_cluster_kernel(const float* __restrict__ scores, long long* __restrict__ indices,
                              float* __restrict__ values, int row_len, int k) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ unsigned hist[256];
  __shared__ unsigned s_merged[256];
  __shared__ unsigned s_prefix, s_want, s_count_gt, s_eq_count, s_tie_base;
  __shared__ unsigned s_warp_sum[kTopkThreads / 32];
  __shared__ unsigned long long s_list[kTopkMax];
  extern __shared__ unsigned s_keys[];  // this block's slice of the keys (dynamic: up to kTopkSliceMax)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = cluster.block_rank();
  const int image = blockIdx.x / kTopkCluster;
  const float* row = scores + (long long)image * row_len;
  const int slice = (row_len + kTopkCluster - 1) / kTopkCluster;
  const int lo = (int)rank * slice;
  const int n_mine = max(0, min(row_len, lo + slice) - lo);

  for (int i = tid; i < n_mine; i += 4 * kTopkThreads) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i + u * kTopkThreads < n_mine) ? row[lo + i + u * kTopkThreads] : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * kTopkThreads < n_mine) s_keys[i + u * kTopkThreads] = topk_key(v[u]);
  }
  if (tid == 0) {
    s_prefix = 0u;
    s_want = (unsigned)k;
    s_count_gt = 0u;
    s_eq_count = 0u;
    s_tie_base = 0u;
  }
  for (int e = tid; e < kTopkMax; e += kTopkThreads) s_list[e] = 0ull;
  __syncthreads();

  // ---- radix select over the cluster ---------------------------------------------------------------------
  unsigned prefix_mask = 0u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    for (int i0 = 0; i0 < n_mine; i0 += kTopkThreads) {
      const int i = i0 + tid;
      unsigned bin = 256u;
      if (i < n_mine) {
        const unsigned key = s_keys[i];
        if ((key & prefix_mask) == prefix) bin = (key >> shift) & 255u;
      }
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin < 256u && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned)__popc(peers));
    }
    cluster.sync();  // all eight tables are complete
    if (tid < 256) {
      unsigned sum = 0u;
#pragma unroll
      for (int r = 0; r < kTopkCluster; ++r) sum += cluster.map_shared_rank(hist, r)[tid];
      s_merged[tid] = sum;
    }
    __syncthreads();
    if (warp == 0) {
      unsigned h[8], mine = 0u;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        h[b] = s_merged[8 * lane + b];
        mine += h[b];
      }
      unsigned suffix = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += t;
      }
      const unsigned above = suffix - mine;
      const unsigned want = s_want;
      __syncwarp();  // every lane has read s_want before the owning lane rewrites it
      if (above < want && want <= suffix) {
        unsigned cum = above, cum_at = 0u;
        int chosen = -1;
#pragma unroll
        for (int b = 7; b >= 0; --b) {
          if (chosen < 0) {
            if (want <= cum + h[b]) {
              chosen = b;
              cum_at = cum;
            } else {
              cum += h[b];
            }
          }
        }
        s_prefix = prefix | ((unsigned)(8 * lane + chosen) << shift);
        s_want = want - cum_at;
      }
    }
    prefix_mask |= 255u << shift;
    cluster.sync();  // every block has read every table (and published its own decision) before the next zero-fill
  }
  const unsigned T = s_prefix, ties = s_want, n_gt = (unsigned)k - ties;

  // ---- collection into block 0's list ---------------------------------------------------------------------
  unsigned long long* list0 = cluster.map_shared_rank(s_list, 0);
  unsigned* count0 = cluster.map_shared_rank(&s_count_gt, 0);
  unsigned my_eq = 0u;
  for (int i0 = 0; i0 < n_mine; i0 += kTopkThreads) {
    const int i = i0 + tid;
    unsigned key = 0u;
    bool gt = false, eq = false;
    if (i < n_mine) {
      key = s_keys[i];
      gt = key > T;
      eq = key == T;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, gt);
    if (bal) {
      unsigned base = 0u;
      if (lane == 0) base = atomicAdd(count0, (unsigned)__popc(bal));  // one remote atomic per warp
      base = __shfl_sync(0xffffffffu, base, 0);
      if (gt) list0[base + __popc(bal & ((1u << lane) - 1u))] =
                  ((unsigned long long)key << 32) | (0xffffffffu - (unsigned)(lo + i));
    }
    my_eq += eq ? 1u : 0u;
  }
  // ties at the threshold: the first `ties` elements equal to T in index order, i.e. lower ranks first
#pragma unroll
  for (int o = 16; o; o >>= 1) my_eq += __shfl_xor_sync(0xffffffffu, my_eq, o);
  if (lane == 0 && my_eq) atomicAdd(&s_eq_count, my_eq);
  cluster.sync();
  unsigned eq_before = 0u;
  for (unsigned r = 0; r < rank; ++r) eq_before += *cluster.map_shared_rank(&s_eq_count, r);
  if (s_eq_count != 0u && eq_before < ties) {  // block-uniform
    for (int i0 = 0; i0 < n_mine; i0 += kTopkThreads) {
      const int i = i0 + tid;
      const bool eq = i < n_mine && s_keys[i] == T;
      const int chunk_ties = __syncthreads_count(eq);
      if (chunk_ties == 0) continue;
      const unsigned bal = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s_warp_sum[warp] = __popc(bal);
      __syncthreads();
      unsigned base = eq_before + s_tie_base;
      for (int w = 0; w < warp; ++w) base += s_warp_sum[w];
      if (eq) {
        const unsigned rnk = base + __popc(bal & ((1u << lane) - 1u));
        if (rnk < ties) list0[n_gt + rnk] = ((unsigned long long)T << 32) | (0xffffffffu - (unsigned)(lo + i));
      }
      __syncthreads();
      if (tid == 0) s_tie_base += (unsigned)chunk_ties;
      __syncthreads();
    }
  }
  cluster.sync();  // block 0's list is complete; the other blocks are done
  if (rank != 0) return;

  unsigned long long v = s_list[tid];
  for (int size = 2; size <= kTopkMax; size <<= 1) {
    const bool desc = (tid & size) == 0;
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      unsigned long long other;
      if (stride >= 32) {
        s_list[tid] = v;
        __syncthreads();
        other = s_list[tid ^ stride];
        __syncthreads();
      } else {
        other = __shfl_xor_sync(0xffffffffu, v, stride);
      }
      const bool keep_max = ((tid & stride) == 0) == desc;
      v = keep_max ? (v > other ? v : other) : (v < other ? v : other);
    }
  }
  if (tid < k) {
    const unsigned idx = 0xffffffffu - (unsigned)(v & 0xffffffffull);
    indices[(long long)image * k + tid] = (long long)idx;
    if (values != nullptr) values[(long long)image * k + tid] = row[idx];
  }
}
}  // namespace

extern "C" {

int msda_rowmax_f32(msda_stream_t stream, const float* logits, long long rows, int num_classes, float* scores) {
  if (rows < 0 || num_classes < 1)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "rows=%lld num_classes=%d", rows, num_classes);
  if (rows == 0) return MSDA_OK;
  if (!logits || !scores) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (num_classes >= 512)
    msda_rowmax_kernel<32><<<aux_grid(rows), kAuxThreads, 0, (cudaStream_t)stream>>>(logits, scores, rows, num_classes);
  else
    msda_rowmax_kernel<8><<<aux_grid((rows + 3) / 4), kAuxThreads, 0, (cudaStream_t)stream>>>(logits, scores, rows,
                                                                                            num_classes);
  return after_launch("msda_rowmax_kernel");
}

int msda_topk_rows_f32(msda_stream_t stream, const float* scores, int batch, int row_len, int k, int64_t* indices,
                       float* values) {
  if (batch < 0 || row_len < 1 || k < 1 || k > row_len)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "batch=%d row_len=%d k=%d", batch, row_len, k);
  if (k > kTopkMax) return fail(MSDA_ERR_UNSUPPORTED, "k=%d: at most %d winners per row", k, kTopkMax);
  if (batch == 0) return MSDA_OK;
  if (!scores || !indices) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 indices");
  // default: a cluster of 8 blocks per image (rows of up to 98,304 scores); MSDA_TOPK_NO_CLUSTER=1 or longer rows:
  // one block per image
  static const bool no_cluster = getenv("MSDA_TOPK_NO_CLUSTER") && atoi(getenv("MSDA_TOPK_NO_CLUSTER")) != 0;
  if (!no_cluster && (row_len + kTopkCluster - 1) / kTopkCluster <= kTopkSliceMax) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)batch * kTopkCluster);
    cfg.blockDim = dim3(kTopkThreads);
    cfg.dynamicSmemBytes = (size_t)((row_len + kTopkCluster - 1) / kTopkCluster) * sizeof(unsigned);
    cfg.stream = (cudaStream_t)stream;
    if (cfg.dynamicSmemBytes > 32 * 1024) {  // static tables take ~10 KB of the 48 KB that need no opt-in
      const int rc = check_cuda(cudaFuncSetAttribute(msda_topk_rows_cluster_kernel,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     kTopkSliceMax * (int)sizeof(unsigned)),
                                "shared-memory opt-in of msda_topk_rows_cluster_kernel");
      if (rc != MSDA_OK) return rc;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kTopkCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    long long* idx_ll = reinterpret_cast<long long*>(indices);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, msda_topk_rows_cluster_kernel, scores, idx_ll, values, row_len, k);
    if (e != cudaSuccess) return check_cuda(e, "launch of msda_topk_rows_cluster_kernel");
    return after_launch("msda_topk_rows_cluster_kernel");
  }
  // rows of up to 49,152 scores are staged in shared memory as keys (192 KB); longer rows are re-read from L2
  const bool staged = row_len <= kTopkStageMax;
  const size_t dyn = staged ? (size_t)row_len * sizeof(unsigned) : 0;
  if (dyn > 48 * 1024) {
    const int rc = check_cuda(cudaFuncSetAttribute(msda_topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   kTopkStageMax * (int)sizeof(unsigned)),
                              "shared-memory opt-in of msda_topk_rows_kernel");
    if (rc != MSDA_OK) return rc;
  }
  msda_topk_rows_kernel<<<batch, kTopkThreads, dyn, (cudaStream_t)stream>>>(
      scores, reinterpret_cast<long long*>(indices), values, row_len, k, staged);
  return after_launch("msda_topk_rows_kernel");
}

}  // extern "C"
