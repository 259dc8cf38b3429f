// msda_launch_other.cu — any-shape kernels (msda_generic.cuh), deterministic grad_value
// (msda_det.cuh) and the corner-index probe.
#include "msda_host.h"
#include "msda_det.cuh"
#include "msda_generic.cuh"

namespace msda {
namespace {
int generic_grid(const MsdaDims& d) {
  const long long tasks = (long long)d.batch * d.num_query * d.num_heads;
  long long blocks = (tasks + 7) / 8;
  if (blocks > 148ll * 64) blocks = 148ll * 64;
  return (int)(blocks < 1 ? 1 : blocks);
}
}  // namespace

template <typename TV, typename TA>
int fwd_generic(cudaStream_t s, const Problem& pb, const TV* value, const TA* loc, const TA* attw, TV* out) {
  msda_fwd_generic_kernel<TV, TA><<<generic_grid(pb.d), 256, 0, s>>>(value, loc, attw, out, pb.lv, pb.d);
  return after_launch("msda_fwd_generic_kernel");
}

template <typename TV, typename TA>
int bwd_generic(cudaStream_t s, const Problem& pb, bool scatter, const TV* go, const TV* value, const TA* loc,
                const TA* attw, TA* gv, TA* gl, TA* ga) {
  if (scatter)
    msda_bwd_generic_kernel<TV, TA, true><<<generic_grid(pb.d), 256, 0, s>>>(go, value, loc, attw, gv, gl, ga, pb.lv, pb.d);
  else
    msda_bwd_generic_kernel<TV, TA, false><<<generic_grid(pb.d), 256, 0, s>>>(go, value, loc, attw, gv, gl, ga, pb.lv, pb.d);
  return after_launch("msda_bwd_generic_kernel");
}

template <typename TV>
int det_grad_value(cudaStream_t s, const Problem& pb, const TV* go, const float* loc, const float* attw,
                   float* gv, void* workspace, size_t workspace_bytes) {
  return deterministic_grad_value<TV>(s, pb.d, pb.lv, go, loc, attw, gv, workspace, workspace_bytes);
}

size_t det_workspace_bytes(int batch, int spatial_size, int num_heads, int channels, int num_levels,
                           int num_query, int num_point) {
  // the larger of the two deterministic paths: sort-by-corner ids (msda_det.cuh) and the window kernel's
  // fixed-point accumulators (256 + 8 bytes per grad_value element, msda_launch_win.cu)
  const size_t ids = deterministic_workspace_bytes(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (batch < 1 || spatial_size < 1 || num_heads < 1 || channels < 1) return ids;
  const size_t fixed = 256 + (size_t)batch * spatial_size * num_heads * channels * 8;
  return ids > fixed ? ids : fixed;
}

int corners_probe(cudaStream_t s, const MsdaLevels& lv, const float* loc, int32_t* corners, long long n,
                  int num_levels, int num_point) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  msda_corners_kernel<<<(int)blocks, 256, 0, s>>>(loc, corners, lv, n, num_levels, num_point);
  return after_launch("msda_corners_kernel");
}

#define INST(TV, TA)                                                                                         \
  template int fwd_generic<TV, TA>(cudaStream_t, const Problem&, const TV*, const TA*, const TA*, TV*);      \
  template int bwd_generic<TV, TA>(cudaStream_t, const Problem&, bool, const TV*, const TV*, const TA*,      \
                                   const TA*, TA*, TA*, TA*);
INST(float, float)
INST(double, double)
INST(__nv_bfloat16, float)
#undef INST
template int det_grad_value<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, float*,
                                   void*, size_t);
template int det_grad_value<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const float*,
                                           const float*, float*, void*, size_t);

}  // namespace msda
