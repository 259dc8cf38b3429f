// ref_legacy_shim.cu — exposes the REFERENCE's own CUDA launchers through a C ABI so that the
// legacy kernels (recompiled for sm_100a, unmodified) can be timed and compared on the GPU box.
// TEST / BENCH INFRASTRUCTURE ONLY.  The reference source is compiled from where it lies
// (-I /root/reference/models/richsem/ops/src/cuda); nothing of it is copied into this repo.
// Launchers wrapped: ms_deformable_im2col_cuda / ms_deformable_col2im_cuda
// (models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:923-954, 956-1327).
#include <cstdint>
#include <cuda_runtime.h>

#include "ms_deform_im2col_cuda.cuh"

extern "C" {

void legacy_forward_f32(cudaStream_t stream, const float* value, const int64_t* shapes, const int64_t* start,
                        const float* loc, const float* attw, int batch, int spatial_size, int num_heads,
                        int channels, int num_levels, int num_query, int num_point, float* out) {
  ms_deformable_im2col_cuda<float>(stream, value, shapes, start, loc, attw, batch, spatial_size, num_heads,
                                   channels, num_levels, num_query, num_point, out);
}

void legacy_backward_f32(cudaStream_t stream, const float* grad_out, const float* value, const int64_t* shapes,
                         const int64_t* start, const float* loc, const float* attw, int batch,
                         int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                         int num_point, float* grad_value, float* grad_loc, float* grad_attw) {
  ms_deformable_col2im_cuda<float>(stream, grad_out, value, shapes, start, loc, attw, batch, spatial_size,
                                   num_heads, channels, num_levels, num_query, num_point, grad_value, grad_loc,
                                   grad_attw);
}

void legacy_forward_f64(cudaStream_t stream, const double* value, const int64_t* shapes, const int64_t* start,
                        const double* loc, const double* attw, int batch, int spatial_size, int num_heads,
                        int channels, int num_levels, int num_query, int num_point, double* out) {
  ms_deformable_im2col_cuda<double>(stream, value, shapes, start, loc, attw, batch, spatial_size, num_heads,
                                    channels, num_levels, num_query, num_point, out);
}

}  // extern "C"
