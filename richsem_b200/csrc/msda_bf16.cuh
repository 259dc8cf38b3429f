// msda_bf16.cuh — bf16-value variant: value / out / grad_out are bfloat16, sampling
// locations, attention weights and all three gradients are fp32, arithmetic is fp32.
// No reference counterpart (the reference dispatches float/double only,
// models/richsem/ops/src/cuda/ms_deform_attn_cuda.cu:64); parity target is the fp32
// oracle within 1e-2.
#pragma once

#include <cuda_bf16.h>

#include "msda_common.cuh"
#include "msda_det.cuh"
#include "msda_generic.cuh"
#include "msda_host.h"

namespace msda {

inline int generic_grid_for(const MsdaDims& d) {
  const long long tasks = (long long)d.batch * d.num_query * d.num_heads;
  long long blocks = (tasks + 7) / 8;
  if (blocks > 148ll * 64) blocks = 148ll * 64;
  return (int)(blocks < 1 ? 1 : blocks);
}

inline int forward_bf16(cudaStream_t s, const MsdaDims& d, const MsdaLevels& lv, const int32_t* order,
                        int order_len, uint32_t flags, const uint16_t* value, const float* loc,
                        const float* attw, uint16_t* out) {
  (void)order; (void)order_len; (void)flags;
  msda_fwd_generic_kernel<__nv_bfloat16, float><<<generic_grid_for(d), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(value), loc, attw, reinterpret_cast<__nv_bfloat16*>(out), lv, d);
  return after_launch("msda_fwd_generic_kernel<bf16>");
}

inline int backward_bf16(cudaStream_t s, const MsdaDims& d, const MsdaLevels& lv, const int32_t* order,
                         int order_len, uint32_t flags, const uint16_t* grad_out, const uint16_t* value,
                         const float* loc, const float* attw, float* gv, float* gl, float* ga, void* ws,
                         size_t ws_bytes) {
  (void)order; (void)order_len;
  const __nv_bfloat16* go = reinterpret_cast<const __nv_bfloat16*>(grad_out);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(value);
  if (flags & MSDA_FLAG_DETERMINISTIC) {
    msda_bwd_generic_kernel<__nv_bfloat16, float, false><<<generic_grid_for(d), 256, 0, s>>>(go, v, loc, attw, gv, gl, ga, lv, d);
    int rc = after_launch("msda_bwd_generic_kernel<bf16,noscatter>");
    if (rc) return rc;
    return deterministic_grad_value<__nv_bfloat16>(s, d, lv, go, loc, attw, gv, ws, ws_bytes);
  }
  msda_bwd_generic_kernel<__nv_bfloat16, float, true><<<generic_grid_for(d), 256, 0, s>>>(go, v, loc, attw, gv, gl, ga, lv, d);
  return after_launch("msda_bwd_generic_kernel<bf16>");
}

}  // namespace msda
