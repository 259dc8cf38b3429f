// msda_host.h — host-side declarations shared by the translation units of libmsda_b200.so.
// msda_capi.cu holds the C entry points, validation and kernel selection; the kernels are instantiated
// and launched from msda_launch_{d32,win,other}.cu so that nvcc can compile them in parallel.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "msda_common.cuh"

namespace msda {

// Records a printf-style message for msda_last_error() on this thread and returns `code`.
int fail(int code, const char* fmt, ...);
// MSDA_OK, or MSDA_ERR_CUDA with the runtime's message.
int check_cuda(cudaError_t e, const char* what);
// Counts the launch (msda_launch_count) and reports a launch-time error, if any.
int after_launch(const char* what);

// One validated call: dimensions, resolved level table, optional query order, flags.
struct Problem {
  MsdaDims d;
  MsdaLevels lv;
  const int32_t* order;
  int order_len;
  uint32_t flags;
  int kernel_hint;  // MSDA_KERNEL_*
  MsdaFused fz;  // ref_dim 0: the reference op; 2 | 4: fused prologue (loc / attw are the raw Linear outputs)
  // the library's own zero-fill kernel of grad_value is the launch right before the backward kernel on this stream: the
  // backward may start early (programmatic dependent launch) and waits (griddepcontrol.wait) before its first reduction
  bool pdl_after_fill;
};

// Kernel family of a head_dim-32 problem (MSDA_KERNEL_*).  Few (query, head) pairs (decoder cross-attention): one
// warp per pair (split).  Large problems: the backward's shared-memory window kernel needs spatially coherent
// tiles of 64 queries, i.e. the host's patch order (encoder self-attention); without an order the samples of a
// tile scatter over the whole map and the tiled L1-gather kernel is the better fit.
inline int kernel_family(const Problem& pb) {
  if (pb.kernel_hint == MSDA_KERNEL_SPLIT) return MSDA_KERNEL_SPLIT;
  if (pb.kernel_hint == MSDA_KERNEL_TILED || pb.kernel_hint == MSDA_KERNEL_WINDOW) return pb.kernel_hint;
  if ((long long)pb.d.batch * pb.d.num_query * pb.d.num_heads <= 65536) return MSDA_KERNEL_SPLIT;
  return pb.order ? MSDA_KERNEL_WINDOW : MSDA_KERNEL_TILED;
}
inline bool use_split(const Problem& pb) { return kernel_family(pb) == MSDA_KERNEL_SPLIT; }

// Zero-fill of grad_value as a kernel (16-byte stores) that lets the next kernel on the stream launch early.
int zero_fill_pdl(cudaStream_t s, float* p, size_t bytes);

// ---- head_dim 32, 4 points, 3..5 levels; VT = float | __nv_bfloat16 (explicitly instantiated) ----
// msda_launch_d32.cu: L1-gather kernels (tiled for large problems, split for small ones).
template <typename VT>
int fwd_d32(cudaStream_t s, const Problem& pb, const VT* value, const float* loc, const float* attw, VT* out);
template <typename VT, bool kScatter>
int bwd_d32(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
            const float* attw, float* gv, float* gl, float* ga);
// msda_launch_win.cu: shared-memory window kernels (large problems).
template <typename VT>
int bwd_d32_win(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
                const float* attw, float* gv, float* gl, float* ga);

// deterministic variant of the window backward (large problems): canonical order inside a block, fixed-point
// accumulation across blocks; needs 256 + 8 bytes per grad_value element of workspace
template <typename VT>
int bwd_d32_win_det(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
                    const float* attw, float* gv, float* gl, float* ga, void* workspace, size_t workspace_bytes);
// ---- msda_launch_other.cu: any-shape kernels, deterministic grad_value, index probe ----
template <typename TV, typename TA>
int fwd_generic(cudaStream_t s, const Problem& pb, const TV* value, const TA* loc, const TA* attw, TV* out);
template <typename TV, typename TA>
int bwd_generic(cudaStream_t s, const Problem& pb, bool scatter, const TV* go, const TV* value, const TA* loc,
                const TA* attw, TA* gv, TA* gl, TA* ga);
template <typename TV>
int det_grad_value(cudaStream_t s, const Problem& pb, const TV* go, const float* loc, const float* attw,
                   float* gv, void* workspace, size_t workspace_bytes);
size_t det_workspace_bytes(int batch, int spatial_size, int num_heads, int channels, int num_levels,
                           int num_query, int num_point);
int corners_probe(cudaStream_t s, const MsdaLevels& lv, const float* loc, int32_t* corners, long long n,
                  int num_levels, int num_point);

}  // namespace msda
