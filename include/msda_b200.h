/*
 * msda_b200.h — C ABI of libmsda_b200.so: multi-scale deformable attention
 * (MSDeformAttn) forward / backward for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for RichSem's native extension
 * `MultiScaleDeformableAttention` (reference: models/richsem/ops/src/).  Every
 * entry point replaces one templated launcher of the reference and keeps its
 * argument order; the trailing `const msda_opts*` is the only addition and may
 * be NULL.
 *
 *   msda_forward_{f32,f64,bf16}   <-  ms_deformable_im2col_cuda<scalar_t>
 *                                     (models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:923-954)
 *   msda_backward_{f32,f64,bf16}  <-  ms_deformable_col2im_cuda<scalar_t>
 *                                     (models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:956-1327)
 *
 * Tensor layouts (all contiguous, exactly the reference's,
 * models/richsem/ops/src/cuda/ms_deform_attn_cuda.cu:40-48):
 *   value              [batch, spatial_size, num_heads, channels]
 *   spatial_shapes     [num_levels, 2]  int64, rows are (H_l, W_l)
 *   level_start_index  [num_levels]     int64
 *   sampling_loc       [batch, num_query, num_heads, num_levels, num_point, 2]  (x, y) in [0,1]
 *   attn_weight        [batch, num_query, num_heads, num_levels, num_point]
 *   out / grad_out     [batch, num_query, num_heads * channels]
 *
 * Differences from the reference launchers, all deliberate:
 *   - every function returns a status (the reference printf()s launch errors
 *     and carries on, cuh:948-952, 1321-1325);
 *   - one launch covers the whole batch: `im2col_step` chunking
 *     (ms_deform_attn_cuda.cu:50-75) is a host-side loop over independent
 *     images and is validated by the host wrapper only;
 *   - the backward zero-fills grad_value itself (the reference relies on
 *     at::zeros_like, ms_deform_attn_cuda.cu:121-123) unless
 *     MSDA_FLAG_GRAD_VALUE_PREZEROED is set; grad_sampling_loc and
 *     grad_attn_weight are fully overwritten;
 *   - the per-level table travels to the kernels in kernel-parameter constant
 *     memory, so the library needs the table on the host.  `spatial_shapes`
 *     and `level_start_index` stay DEVICE pointers as in the reference; pass a
 *     host mirror in msda_opts to avoid a blocking device->host copy (required
 *     under CUDA-graph capture).
 *
 * No function allocates device memory, synchronises the device (except the
 * mirror-less path noted above) or retains any pointer after it returns.
 * All work is enqueued on `stream`.  Thread-safe and re-entrant.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 3
#define MSDA_MAX_LEVELS 16

/* status codes */
#define MSDA_OK 0
#define MSDA_ERR_INVALID_ARGUMENT 1 /* bad pointer / dimension / alignment        */
#define MSDA_ERR_UNSUPPORTED 2      /* shape outside what the library implements */
#define MSDA_ERR_CUDA 3             /* a CUDA runtime call or launch failed       */
#define MSDA_ERR_WORKSPACE 4        /* deterministic mode: workspace missing/small */

/* msda_opts.flags */
#define MSDA_FLAG_DETERMINISTIC 0x1u          /* bitwise reproducible grad_value (fixed-point accumulation of canonically
                                                 ordered partial sums, or sort-by-corner segmented sums); needs workspace */
#define MSDA_FLAG_GRAD_VALUE_PREZEROED 0x2u   /* caller already zeroed grad_value (or wants accumulation)           */
#define MSDA_FLAG_FORCE_GENERIC 0x4u          /* bypass the head_dim-32 kernels (testing)                           */
#define MSDA_FLAG_COORDS_FMA 0x10u            /* pixel coordinate = fma(loc, size, -0.5): what nvcc -fmad=true makes of
                                                 cuh:285-286, i.e. the compiled reference; default is mul-then-sub  */
#define MSDA_FLAG_NO_GRAD_VALUE 0x400u        /* backward: value needs no gradient; grad_value is not touched and may be NULL */

/* msda_opts.kernel_hint: which kernel family serves a head_dim-32 problem.  0 lets the library decide (split for
 * at most 65,536 (query, head) pairs, else window when a query order is given, else tiled); the other values
 * exist so that tests and tuning runs can put a small problem through the large-problem kernels. */
#define MSDA_KERNEL_AUTO 0
#define MSDA_KERNEL_SPLIT 1  /* one warp per (query, head): decoder-sized problems                                  */
#define MSDA_KERNEL_TILED 2  /* one lane group per (query, head), L1 gather, one L2 reduction per sampled corner    */
#define MSDA_KERNEL_WINDOW 3 /* backward: shared-memory window, cell-sorted on-chip merging (forward: tiled)        */

typedef void* msda_stream_t; /* cudaStream_t */

typedef struct msda_opts {
  uint32_t struct_size;                   /* sizeof(msda_opts); lets the struct grow */
  uint32_t flags;
  const int64_t* spatial_shapes_host;     /* optional host mirror of spatial_shapes [num_levels*2]   */
  const int64_t* level_start_index_host;  /* optional host mirror of level_start_index [num_levels]  */
  /* Optional processing order of the queries (DEVICE int32 array).  Entry i is a
   * query index in [0, num_query) or -1 (skip).  Every query must appear exactly
   * once.  It only changes which thread block handles which query, i.e. cache
   * locality, never results.  NULL = natural order. */
  const int32_t* query_order;
  int32_t query_order_len;
  int32_t kernel_hint;                    /* MSDA_KERNEL_*; 0 = automatic */
  void* workspace;                        /* DEVICE scratch for MSDA_FLAG_DETERMINISTIC */
  size_t workspace_bytes;
} msda_opts;

/* ---- forward --------------------------------------------------------------
 * out[b,q,m,:] = sum_{l,p} attn_weight[b,q,m,l,p] * bilinear(value_l[b,:,m,:]; x*W_l-0.5, y*H_l-0.5)
 * zero padding; semantics of cuh:237-299 + cuh:33-84. */
int msda_forward_f32(msda_stream_t stream, const float* value, const int64_t* spatial_shapes,
                     const int64_t* level_start_index, const float* sampling_loc,
                     const float* attn_weight, int batch, int spatial_size, int num_heads,
                     int channels, int num_levels, int num_query, int num_point, float* out,
                     const msda_opts* opts);
int msda_forward_f64(msda_stream_t stream, const double* value, const int64_t* spatial_shapes,
                     const int64_t* level_start_index, const double* sampling_loc,
                     const double* attn_weight, int batch, int spatial_size, int num_heads,
                     int channels, int num_levels, int num_query, int num_point, double* out,
                     const msda_opts* opts);
/* bf16 variant (new capability, no reference counterpart): value and out are
 * bfloat16 bit patterns (uint16_t); sampling_loc / attn_weight stay fp32;
 * accumulation is fp32. */
int msda_forward_bf16(msda_stream_t stream, const uint16_t* value, const int64_t* spatial_shapes,
                      const int64_t* level_start_index, const float* sampling_loc,
                      const float* attn_weight, int batch, int spatial_size, int num_heads,
                      int channels, int num_levels, int num_query, int num_point, uint16_t* out,
                      const msda_opts* opts);

/* ---- backward -------------------------------------------------------------
 * gradient formulas of cuh:87-159; argument order of cuh:956-973. */
int msda_backward_f32(msda_stream_t stream, const float* grad_out, const float* value,
                      const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const float* sampling_loc, const float* attn_weight, int batch,
                      int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, float* grad_value, float* grad_sampling_loc,
                      float* grad_attn_weight, const msda_opts* opts);
int msda_backward_f64(msda_stream_t stream, const double* grad_out, const double* value,
                      const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const double* sampling_loc, const double* attn_weight, int batch,
                      int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, double* grad_value, double* grad_sampling_loc,
                      double* grad_attn_weight, const msda_opts* opts);
/* bf16 variant: grad_out and value are bfloat16; all three gradients are fp32. */
int msda_backward_bf16(msda_stream_t stream, const uint16_t* grad_out, const uint16_t* value,
                       const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const float* sampling_loc, const float* attn_weight, int batch,
                       int spatial_size, int num_heads, int channels, int num_levels,
                       int num_query, int num_point, float* grad_value, float* grad_sampling_loc,
                       float* grad_attn_weight, const msda_opts* opts);

/* ---- fused prologue ---------------------------------
 * The module computes sampling locations and attention weights from the raw outputs of two Linear layers
 * (models/richsem/ops/modules/ms_deform_attn.py:98-111: softmax over L*P, reference point + normalised
 * offset) in five elementwise PyTorch kernels around the op.  These entry points take the RAW tensors
 *   sampling_offsets  [batch, num_query, num_heads, num_levels, num_point, 2]
 *   attn_logits       [batch, num_query, num_heads, num_levels * num_point]
 *   reference_points  [batch, num_query, num_levels, ref_dim]      ref_dim 2: points, 4: (cx, cy, w, h) boxes
 * and do that arithmetic inside the kernels' decode step (same operations in the same order, so corner
 * indices are unchanged); the backward returns the gradients of the raw tensors (no gradient for
 * reference_points).  Implemented for large head_dim-32 / 4-point problems with the default kernels;
 * otherwise MSDA_ERR_UNSUPPORTED and the caller uses the unfused entry points. */
int msda_forward_fused_f32(msda_stream_t stream, const float* value, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const float* sampling_offsets,
                           const float* attn_logits, const float* reference_points, int ref_dim, int batch,
                           int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                           int num_point, float* out, const msda_opts* opts);
int msda_forward_fused_bf16(msda_stream_t stream, const uint16_t* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const float* sampling_offsets,
                            const float* attn_logits, const float* reference_points, int ref_dim, int batch,
                            int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                            int num_point, uint16_t* out, const msda_opts* opts);
int msda_backward_fused_f32(msda_stream_t stream, const float* grad_out, const float* value,
                            const int64_t* spatial_shapes, const int64_t* level_start_index,
                            const float* sampling_offsets, const float* attn_logits,
                            const float* reference_points, int ref_dim, int batch, int spatial_size,
                            int num_heads, int channels, int num_levels, int num_query, int num_point,
                            float* grad_value, float* grad_sampling_offsets, float* grad_attn_logits,
                            const msda_opts* opts);
int msda_backward_fused_bf16(msda_stream_t stream, const uint16_t* grad_out, const uint16_t* value,
                             const int64_t* spatial_shapes, const int64_t* level_start_index,
                             const float* sampling_offsets, const float* attn_logits,
                             const float* reference_points, int ref_dim, int batch, int spatial_size,
                             int num_heads, int channels, int num_levels, int num_query, int num_point,
                             float* grad_value, float* grad_sampling_offsets, float* grad_attn_logits,
                             const msda_opts* opts);

/* ---- value preparation --------------------------------
 * The module zeroes the padded tokens of the projected value (models/richsem/ops/modules/ms_deform_attn.py:95-96,
 * `value.masked_fill(input_padding_mask[..., None], 0)`) before it is viewed as [batch, spatial_size, heads, channels]
 * (:97).  `rows` = batch * spatial_size, `row_elems` = heads * channels (a multiple of 4); padding_mask is the
 * module's bool mask, one byte per token, non-zero = padding.
 *   msda_value_prepare_bf16    value_out[row,:] = mask[row] ? 0 : bf16(projected[row,:])   (mask may be NULL: cast only);
 *                              feeds the bf16 kernels in one pass instead of masked_fill + cast
 *   msda_zero_masked_rows_f32  in place: data[row,:] = 0 where mask[row]; only masked rows are written.  Serves the
 *                              fp32 forward (on the projection's output) and the backward of both (on grad_value). */
int msda_value_prepare_bf16(msda_stream_t stream, const float* projected, const uint8_t* padding_mask,
                            long long rows, int row_elems, uint16_t* value_out);
int msda_zero_masked_rows_f32(msda_stream_t stream, float* data, const uint8_t* padding_mask, long long rows,
                              int row_elems);

/* ---- two-stage proposals ------------------------------
 * gen_encoder_output_proposals (models/richsem/utils.py:10-65) in one pass over the encoder memory:
 *   memory            [batch, spatial_size, channels] fp32 (channels a multiple of 4)
 *   padding_mask      [batch, spatial_size] bytes, non-zero = padding; NULL = no padding
 *   spatial_shapes    [num_levels, 2] int64 DEVICE (H_l, W_l); levels lie back to back (utils.py:25,54); host mirror
 *                     through msda_opts.spatial_shapes_host as for the op
 *   wh_base           DEVICE float[2] = sigmoid(learnedwh) (utils.py:41), or NULL for the constant 0.05 (:43)
 *   output_memory     [batch, spatial_size, channels]: memory, zero on padded or invalid tokens (:54-56)
 *   output_proposals  [batch, spatial_size, 4]: logit of (cx, cy, w, h), +inf on padded or invalid tokens (:49-52)
 *   valid_hw_workspace  DEVICE int32[batch * num_levels * 2] for the valid H / W (:27-28); only needed with a
 *                     padding mask and batch * num_levels > 128 (smaller tables are recomputed inside the kernel)
 * The backward routes grad_output_memory to the kept tokens (those whose proposal is finite). */
int msda_encoder_proposals_f32(msda_stream_t stream, const float* memory, const uint8_t* padding_mask,
                               const int64_t* spatial_shapes, const float* wh_base, int batch, int spatial_size,
                               int channels, int num_levels, float* output_memory, float* output_proposals,
                               int32_t* valid_hw_workspace, const msda_opts* opts);
int msda_encoder_proposals_backward_f32(msda_stream_t stream, const float* grad_output_memory,
                                        const float* output_proposals, int batch, int spatial_size, int channels,
                                        float* grad_memory);

/* ---- two-stage query selection (second half) -----------
 * models/richsem/deformable_transformer.py:367-369:
 *   topk_proposals = torch.topk(enc_outputs_class_unselected.max(-1)[0], num_queries, dim=1)[1]
 * msda_rowmax_f32     scores[row] = max_c logits[row, c]   (logits [rows, num_classes]; a NaN in a row gives NaN)
 * msda_topk_rows_f32  indices[b, :] = positions of the k largest scores of row b of scores [batch, row_len], sorted by
 *                     descending score; equal scores: lower index first; NaN ranks above everything (torch.topk).
 *                     indices int64 [batch, k]; values (optional, may be NULL) fp32 [batch, k].  k <= 1024, else
 *                     MSDA_ERR_UNSUPPORTED. */
int msda_rowmax_f32(msda_stream_t stream, const float* logits, long long rows, int num_classes, float* scores);
int msda_topk_rows_f32(msda_stream_t stream, const float* scores, int batch, int row_len, int k, int64_t* indices,
                       float* values);

/* ---- encoder-layer epilogue: residual add + LayerNorm ----
 * models/richsem/deformable_transformer.py:871-872 (`src = src + self.dropout1(src2); src = self.norm1(src)`) and
 * :866-867 (the same around the FFN), dropout 0 as RichSem trains (config/RichSem/baseline_4scale.py:42):
 *   out[row,:] = (y - mean(y)) * rstd(y) * gamma + beta,   y = x[row,:] + residual[row,:]
 * x / residual / out / grad_* are [rows, channels] fp32, channels a multiple of 128 up to 512 (else
 * MSDA_ERR_UNSUPPORTED); residual may be NULL (plain LayerNorm), gamma / beta may be NULL (1 / 0); biased variance,
 * rstd = 1 / sqrt(var + eps) as torch.nn.LayerNorm.  mean / rstd [rows] are the statistics the backward wants (both
 * NULL: not stored).
 * The backward writes grad_in = d loss / d y — the gradient of x AND of residual — and, when grad_gamma / grad_beta are
 * given, their sums over the rows in a fixed order (bitwise reproducible), through a caller-provided DEVICE workspace
 * of msda_add_layernorm_workspace_bytes(rows, channels). */
int msda_add_layernorm_f32(msda_stream_t stream, const float* x, const float* residual, const float* gamma,
                           const float* beta, long long rows, int channels, float eps, float* out, float* mean,
                           float* rstd);
int msda_add_layernorm_backward_f32(msda_stream_t stream, const float* grad_out, const float* x, const float* residual,
                                    const float* gamma, const float* mean, const float* rstd, long long rows,
                                    int channels, float* grad_in, float* grad_gamma, float* grad_beta, void* workspace,
                                    size_t workspace_bytes);
size_t msda_add_layernorm_workspace_bytes(long long rows, int channels);

/* ---- index contract probe --------------------------------------------------
 * Writes, for every sample (b,q,m,l,p), the four bilinear corner token indices
 * (level_start_index[l] + h*W_l + w, i.e. an index into the spatial_size axis)
 * in the order (h0,w0) (h0,w1) (h1,w0) (h1,w1); -1 for a corner that
 * contributes nothing (out of the map, or the whole sample skipped by the range
 * test of cuh:288).  corners: int32 [batch,num_query,num_heads,num_levels,num_point,4].
 * Uses the same device function as the production fp32/bf16 kernels. */
int msda_debug_corners_f32(msda_stream_t stream, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const float* sampling_loc, int batch,
                           int num_heads, int num_levels, int num_query, int num_point,
                           int32_t* corners, const msda_opts* opts);

/* ---- misc ------------------------------------------------------------------ */
/* Bytes of device workspace MSDA_FLAG_DETERMINISTIC needs for this problem. */
size_t msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels,
                                     int num_levels, int num_query, int num_point);
/* 1 if (dtype_bytes, channels, num_levels, num_point) is served by the tuned kernels, else 0
 * (the generic kernels serve everything else). dtype_bytes: 4 = f32, 8 = f64, 2 = bf16. */
int msda_has_fast_path(int dtype_bytes, int channels, int num_levels, int num_point);
int msda_abi_version(void);
/* Static build description: "msda_b200 <abi> sm_100a <date>". */
const char* msda_build_info(void);
/* Message for the last non-OK status returned on the calling thread. */
const char* msda_last_error(void);
/* Number of kernels this library has launched since load (all threads); used by bench.py
 * for its gpu_launches figure. */
uint64_t msda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
