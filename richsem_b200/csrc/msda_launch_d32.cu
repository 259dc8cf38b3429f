// msda_launch_d32.cu — instantiates and launches the L1-gather D=32 kernels (msda_d32.cuh):
// tiled and split.
#include "msda_host.h"
#include "msda_d32.cuh"

namespace msda {
namespace {

// ---- tuned D=32 dispatch (fp32 and bf16 value) -----------------------------------------------
template <typename VT, int kL, int kM>
int launch_fwd_d32(cudaStream_t s, const Problem& pb, const VT* value, const float* loc,
                   const float* attw, VT* out) {
  using Cfg = msda::D32Cfg<VT, kL * 4>;
  const int tiles = (pb.order_len + msda::kTileQ - 1) / msda::kTileQ;
  dim3 grid(tiles * pb.d.num_heads, pb.d.batch);
  msda::msda_fwd_d32_kernel<VT, kL, 4, kM><<<grid, msda::kThreads, Cfg::SMEM_BYTES, s>>>(
      value, loc, attw, out, pb.order, pb.order_len, pb.lv, pb.d.spatial_size, pb.d.num_heads,
      pb.d.num_query, pb.fz);
  return after_launch("msda_fwd_d32_kernel");
}
template <typename VT, int kL, int kM, bool kScatter>
int launch_bwd_d32(cudaStream_t s, const Problem& pb, const VT* grad_out, const VT* value,
                   const float* loc, const float* attw, float* gv, float* gl, float* ga) {
  using Cfg = msda::D32Cfg<VT, kL * 4>;
  const int tiles = (pb.order_len + msda::kTileQ - 1) / msda::kTileQ;
  dim3 grid(tiles * pb.d.num_heads, pb.d.batch);
  msda::msda_bwd_d32_kernel<VT, kL, 4, kM, kScatter><<<grid, msda::kThreads, Cfg::SMEM_BYTES, s>>>(
      grad_out, value, loc, attw, gv, gl, ga, pb.order, pb.order_len, pb.lv, pb.d.spatial_size,
      pb.d.num_heads, pb.d.num_query);
  return after_launch("msda_bwd_d32_kernel");
}

template <typename VT, int kL, int kM>
int launch_fwd_split(cudaStream_t s, const Problem& pb, const VT* value, const float* loc,
                     const float* attw, VT* out) {
  constexpr int QPB = msda::kSplitThreads / 32;
  dim3 grid(((pb.d.num_query + QPB - 1) / QPB) * pb.d.num_heads, pb.d.batch);
  msda::msda_fwd_d32_split_kernel<VT, kL, 4, kM><<<grid, msda::kSplitThreads, 0, s>>>(
      value, loc, attw, out, pb.lv, pb.d.spatial_size, pb.d.num_heads, pb.d.num_query, pb.fz);
  return after_launch("msda_fwd_d32_split_kernel");
}
template <typename VT, int kL, int kM, bool kScatter>
int launch_bwd_split(cudaStream_t s, const Problem& pb, const VT* grad_out, const VT* value,
                     const float* loc, const float* attw, float* gv, float* gl, float* ga) {
  constexpr int QPB = msda::kSplitThreads / 32;
  dim3 grid(((pb.d.num_query + QPB - 1) / QPB) * pb.d.num_heads, pb.d.batch);
  if (kScatter && pb.pdl_after_fill) {
    // programmatic dependent launch behind msda_zero_fill_kernel (see zero_fill_pdl)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(msda::kSplitThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, msda::msda_bwd_d32_split_kernel<VT, kL, 4, kM, kScatter>, grad_out, value,
                                             loc, attw, gv, gl, ga, pb.lv, pb.d.spatial_size, pb.d.num_heads,
                                             pb.d.num_query, pb.fz);
    if (e != cudaSuccess) return check_cuda(e, "launch of msda_bwd_d32_split_kernel");
    return after_launch("msda_bwd_d32_split_kernel");
  }
  msda::msda_bwd_d32_split_kernel<VT, kL, 4, kM, kScatter><<<grid, msda::kSplitThreads, 0, s>>>(
      grad_out, value, loc, attw, gv, gl, ga, pb.lv, pb.d.spatial_size, pb.d.num_heads, pb.d.num_query, pb.fz);
  return after_launch("msda_bwd_d32_split_kernel");
}


#define MSDA_SWITCH_L(L_, CALL)                                                              \
  switch (L_) {                                                                              \
    case 3: return CALL(3);                                                                  \
    case 4: return CALL(4);                                                                  \
    case 5: return CALL(5);                                                                  \
    default: return fail(MSDA_ERR_UNSUPPORTED, "no tuned kernel for num_levels=%d", L_);      \
  }

}  // namespace

template <typename VT>
int fwd_d32(cudaStream_t s, const Problem& pb, const VT* value, const float* loc, const float* attw,
            VT* out) {
  if (use_split(pb)) {
    if (pb.d.num_heads == 8 && pb.d.num_levels == 4) return launch_fwd_split<VT, 4, 8>(s, pb, value, loc, attw, out);
#define CALL(L) launch_fwd_split<VT, L, 0>(s, pb, value, loc, attw, out)
    MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
  }
  // the DINO / RichSem configuration (8 heads, 4 levels) gets the head count baked in
  if (pb.d.num_heads == 8 && pb.d.num_levels == 4) return launch_fwd_d32<VT, 4, 8>(s, pb, value, loc, attw, out);
#define CALL(L) launch_fwd_d32<VT, L, 0>(s, pb, value, loc, attw, out)
  MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
}
template <typename VT, bool kScatter>
int bwd_d32(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
            const float* attw, float* gv, float* gl, float* ga) {
  if (use_split(pb)) {
    if (pb.d.num_heads == 8 && pb.d.num_levels == 4)
      return launch_bwd_split<VT, 4, 8, kScatter>(s, pb, go, value, loc, attw, gv, gl, ga);
#define CALL(L) launch_bwd_split<VT, L, 0, kScatter>(s, pb, go, value, loc, attw, gv, gl, ga)
    MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
  }
#define CALL(L) launch_bwd_d32<VT, L, 0, kScatter>(s, pb, go, value, loc, attw, gv, gl, ga)
  MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
}


template int fwd_d32<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, float*);
template int fwd_d32<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const float*, const float*,
                                    __nv_bfloat16*);
#define INST_BWD(VT, SC)                                                                                   \
  template int bwd_d32<VT, SC>(cudaStream_t, const Problem&, const VT*, const VT*, const float*, const float*, \
                               float*, float*, float*);
INST_BWD(float, true)
INST_BWD(float, false)
INST_BWD(__nv_bfloat16, true)
INST_BWD(__nv_bfloat16, false)
#undef INST_BWD

namespace {
// Zero-fill of grad_value.  `griddepcontrol.launch_dependents` right away: the backward kernel launched behind it
// with the programmatic-serialization attribute may become resident as soon as every block of this grid has started,
// and runs its decode / gather / dot-product phase while the fill is still writing.
__global__ void __launch_bounds__(256) msda_zero_fill_kernel(float4* __restrict__ p, const size_t n16) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) p[i] = z;
}
}  // namespace

int zero_fill_pdl(cudaStream_t s, float* p, size_t bytes) {
  const size_t n16 = bytes / 16;
  // many short-lived blocks (not a persistent grid): SM slots free up while the fill is still running, so that the
  // dependent backward kernel's blocks can move in next to it
  size_t blocks = (n16 + 256 * 16 - 1) / (256 * 16);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 64) blocks = 148 * 64;
  msda_zero_fill_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<float4*>(p), n16);
  return after_launch("msda_zero_fill_kernel");
}

}  // namespace msda
