// msda_capi.cu — extern "C" entry points of libmsda_b200.so (see include/msda_b200.h).
// Validation, level-table resolution, kernel selection and launch.  No torch types.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "msda_host.h"

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

namespace msda {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MSDA_OK;
  return fail(MSDA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}
}  // namespace msda

namespace {
using msda::after_launch;
using msda::check_cuda;
using msda::fail;

inline uint32_t opt_flags(const msda_opts* o) { return o ? o->flags : 0u; }

using msda::Problem;

// Builds the constant-memory level table.  With a host mirror this is pure host work;
// without one it is a blocking device->host copy on `stream` (documented slow path).
int resolve_levels(cudaStream_t stream, const int64_t* shapes_dev, const int64_t* start_dev,
                   int num_levels, int spatial_size, const msda_opts* opts, MsdaLevels* lv) {
  if (num_levels < 1 || num_levels > MSDA_MAX_LEVELS)
    return fail(MSDA_ERR_UNSUPPORTED, "num_levels=%d outside [1,%d]", num_levels, MSDA_MAX_LEVELS);
  int64_t shp[2 * MSDA_MAX_LEVELS], st[MSDA_MAX_LEVELS];
  if (opts && opts->spatial_shapes_host && opts->level_start_index_host) {
    memcpy(shp, opts->spatial_shapes_host, sizeof(int64_t) * 2 * num_levels);
    memcpy(st, opts->level_start_index_host, sizeof(int64_t) * num_levels);
  } else {
    if (!shapes_dev || !start_dev)
      return fail(MSDA_ERR_INVALID_ARGUMENT, "spatial_shapes / level_start_index is NULL");
    cudaError_t e = cudaMemcpyAsync(shp, shapes_dev, sizeof(int64_t) * 2 * num_levels,
                                    cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(st, start_dev, sizeof(int64_t) * num_levels, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return check_cuda(e, "copy of the level table to the host");
  }
  memset(lv, 0, sizeof(*lv));
  for (int l = 0; l < num_levels; ++l) {
    const int64_t H = shp[2 * l], W = shp[2 * l + 1], s = st[l];
    if (H < 1 || W < 1 || s < 0 || H > INT32_MAX || W > INT32_MAX || H * W > INT32_MAX ||
        s + H * W > (int64_t)spatial_size)
      return fail(MSDA_ERR_INVALID_ARGUMENT,
                  "level %d: shape (%lld,%lld) start %lld does not fit spatial_size %d", l,
                  (long long)H, (long long)W, (long long)s, spatial_size);
    lv->H[l] = (int)H;
    lv->W[l] = (int)W;
    lv->start[l] = (int)s;
  }
  return MSDA_OK;
}

int make_problem(cudaStream_t stream, const int64_t* shapes, const int64_t* start, int batch,
                 int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                 int num_point, const msda_opts* opts, Problem* pb) {
  if (batch < 0 || spatial_size < 1 || num_heads < 1 || channels < 1 || num_query < 0 || num_point < 1)
    return fail(MSDA_ERR_INVALID_ARGUMENT,
                "bad dimensions: batch=%d spatial_size=%d num_heads=%d channels=%d num_levels=%d "
                "num_query=%d num_point=%d",
                batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (opts && opts->struct_size != sizeof(msda_opts))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_opts.struct_size=%u, library expects %zu",
                opts->struct_size, sizeof(msda_opts));
  pb->d = MsdaDims{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  pb->fz = MsdaFused{nullptr, 0};
  pb->pdl_after_fill = false;
  pb->flags = opt_flags(opts);
  pb->kernel_hint = opts ? opts->kernel_hint : MSDA_KERNEL_AUTO;
  if (pb->kernel_hint < MSDA_KERNEL_AUTO || pb->kernel_hint > MSDA_KERNEL_WINDOW)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_opts.kernel_hint=%d is not one of MSDA_KERNEL_*", pb->kernel_hint);
  pb->order = opts ? opts->query_order : nullptr;
  pb->order_len = pb->order ? opts->query_order_len : num_query;
  if (pb->order && pb->order_len < num_query)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "query_order_len=%d shorter than num_query=%d",
                pb->order_len, num_query);
  const int rc = resolve_levels(stream, shapes, start, num_levels, spatial_size, opts, &pb->lv);
  pb->lv.coord_fma = (pb->flags & MSDA_FLAG_COORDS_FMA) ? 1 : 0;
  return rc;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

bool fast_shape(int dtype_bytes, int channels, int num_levels, int num_point) {
  return (dtype_bytes == 4 || dtype_bytes == 2) && channels == 32 && num_point == 4 &&
         num_levels >= 3 && num_levels <= 5;  // DINO-family pyramids; anything else takes the generic kernels
}

bool fits_int32(const MsdaDims& d) {
  return (long long)d.spatial_size * d.num_heads * d.channels < (1ll << 31);
}

// ---- templated front ends --------------------------------------------------------------------
// TV: storage type of value / out / grad_out;  TA: type of locations, weights and all gradients.
template <typename TV, typename TA>
int forward_impl(cudaStream_t s, const TV* value, const int64_t* shapes, const int64_t* start,
                 const TA* loc, const TA* attw, int batch, int spatial_size, int num_heads,
                 int channels, int num_levels, int num_query, int num_point, TV* out,
                 const msda_opts* opts) {
  if (!value || !loc || !attw || !out) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  Problem pb;
  int rc = make_problem(s, shapes, start, batch, spatial_size, num_heads, channels, num_levels,
                        num_query, num_point, opts, &pb);
  if (rc != MSDA_OK) return rc;
  if (batch == 0 || num_query == 0) return MSDA_OK;
  if constexpr (sizeof(TA) == 4) {
    if (!(pb.flags & MSDA_FLAG_FORCE_GENERIC) && fast_shape(sizeof(TV), channels, num_levels, num_point) &&
        fits_int32(pb.d) && aligned(value, 16) && aligned(loc, 16) && aligned(attw, 16) && aligned(out, 16))
    {
      return msda::fwd_d32<TV>(s, pb, value, loc, attw, out);
    }
  }
  return msda::fwd_generic<TV, TA>(s, pb, value, loc, attw, out);
}

template <typename TV, typename TA>
int backward_impl(cudaStream_t s, const TV* grad_out, const TV* value, const int64_t* shapes,
                  const int64_t* start, const TA* loc, const TA* attw, int batch, int spatial_size,
                  int num_heads, int channels, int num_levels, int num_query, int num_point,
                  TA* gv, TA* gl, TA* ga, const msda_opts* opts) {
  if (!grad_out || !value || !loc || !attw || !gl || !ga || (!gv && !(opt_flags(opts) & MSDA_FLAG_NO_GRAD_VALUE)))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  Problem pb;
  int rc = make_problem(s, shapes, start, batch, spatial_size, num_heads, channels, num_levels,
                        num_query, num_point, opts, &pb);
  if (rc != MSDA_OK) return rc;
  if (batch == 0) return MSDA_OK;
  const bool no_gv = (pb.flags & MSDA_FLAG_NO_GRAD_VALUE) != 0;  // value needs no gradient: gv may be NULL
  const bool det = (pb.flags & MSDA_FLAG_DETERMINISTIC) != 0 && !no_gv;
  if (det && sizeof(TA) != 4)
    return fail(MSDA_ERR_UNSUPPORTED, "deterministic mode is implemented for fp32 / bf16 only");
  const size_t gv_bytes = (size_t)batch * spatial_size * num_heads * channels * sizeof(TA);
  // deterministic mode writes every grad_value row itself
  if (!(pb.flags & MSDA_FLAG_GRAD_VALUE_PREZEROED) && !(det && num_query > 0) && !no_gv) {
    // decoder-sized fp32-gradient problems (split kernel): the fill is a third of the backward's time, so it is a kernel
    // of this library that the backward kernel overlaps with its gather phase (programmatic dependent launch)
    bool pdl = false;
    if constexpr (sizeof(TA) == 4) {
      static const bool no_pdl = getenv("MSDA_B200_NO_PDL") && atoi(getenv("MSDA_B200_NO_PDL")) != 0;
      // ... and the window kernel of encoder-sized problems, whose whole front end precedes its first reduction
      const int fam = msda::kernel_family(pb);
      const bool split = fam == MSDA_KERNEL_SPLIT, window = fam == MSDA_KERNEL_WINDOW;
      pdl = !no_pdl && num_query > 0 && !det && !(pb.flags & MSDA_FLAG_FORCE_GENERIC) &&
            fast_shape(sizeof(TV), channels, num_levels, num_point) && fits_int32(pb.d) && (split || window) &&
            aligned(value, 16) && aligned(gv, 16) && aligned(loc, 16) && aligned(attw, 16) && aligned(grad_out, 16) &&
            aligned(gl, 16) && aligned(ga, 16) && gv_bytes % 16 == 0;
    }
    if (pdl) {
      rc = msda::zero_fill_pdl(s, reinterpret_cast<float*>(gv), gv_bytes);
      pb.pdl_after_fill = true;
    } else {
      rc = check_cuda(cudaMemsetAsync(gv, 0, gv_bytes, s), "zero-fill of grad_value");
    }
    if (rc != MSDA_OK) return rc;
  }
  if (num_query == 0) return MSDA_OK;
  if constexpr (sizeof(TA) == 4) {
    const bool fast = !(pb.flags & MSDA_FLAG_FORCE_GENERIC) &&
                      fast_shape(sizeof(TV), channels, num_levels, num_point) && fits_int32(pb.d) &&
                      aligned(value, 16) && aligned(gv, 16) && aligned(loc, 16) && aligned(attw, 16) &&
                      aligned(grad_out, 16) && aligned(gl, 16) && aligned(ga, 16);  // 16-byte vector loads / stores / reductions
    if (fast && no_gv) return msda::bwd_d32<TV, false>(s, pb, grad_out, value, loc, attw, gv, gl, ga);
    // deterministic, window-sized problem: the window kernel with canonical in-block order and fixed-point accumulation
    const int fam = msda::kernel_family(pb);
    if (fast && det && fam == MSDA_KERNEL_WINDOW)
      return msda::bwd_d32_win_det<TV>(s, pb, grad_out, value, loc, attw, gv, gl, ga, opts ? opts->workspace : nullptr,
                                       opts ? opts->workspace_bytes : 0);
    if (fast) {
      // encoder-sized problems with a patch order: the shared-memory window kernel; otherwise the L1-gather kernels
      // (split: warp per (query, head); tiled: lane group per (query, head)), which in deterministic mode leave
      // grad_value to the sort-by-corner pass below
      rc = (fam == MSDA_KERNEL_WINDOW && !det) ? msda::bwd_d32_win<TV>(s, pb, grad_out, value, loc, attw, gv, gl, ga)
           : det                                ? msda::bwd_d32<TV, false>(s, pb, grad_out, value, loc, attw, gv, gl, ga)
                                                : msda::bwd_d32<TV, true>(s, pb, grad_out, value, loc, attw, gv, gl, ga);
      if (rc != MSDA_OK || !det) return rc;
    }
    if (det) {
      if (!fast) {
        if ((rc = msda::bwd_generic<TV, TA>(s, pb, false, grad_out, value, loc, attw, gv, gl, ga))) return rc;
      }
      return msda::det_grad_value<TV>(s, pb, grad_out, loc, attw, gv, opts ? opts->workspace : nullptr,
                                      opts ? opts->workspace_bytes : 0);
    }
  }
  return msda::bwd_generic<TV, TA>(s, pb, !no_gv, grad_out, value, loc, attw, gv, gl, ga);
}

// ---- fused prologue: head_dim-32 problems served by the window (encoder self-attention) or the split
// (decoder cross-attention) family ----------------------------------------------------------------------------------
bool fused_supported(const Problem& pb, int dtype_bytes) {
  const uint32_t bad = MSDA_FLAG_DETERMINISTIC | MSDA_FLAG_FORCE_GENERIC | MSDA_FLAG_NO_GRAD_VALUE;
  const int fam = msda::kernel_family(pb);
  return !(pb.flags & bad) && fast_shape(dtype_bytes, pb.d.channels, pb.d.num_levels, pb.d.num_point) &&
         fits_int32(pb.d) && (fam == MSDA_KERNEL_WINDOW || fam == MSDA_KERNEL_SPLIT) && pb.d.batch > 0 && pb.d.num_query > 0;
}

template <typename TV>
int forward_fused_impl(cudaStream_t s, const TV* value, const int64_t* shapes, const int64_t* start,
                       const float* offsets, const float* logits, const float* ref, int ref_dim, int batch,
                       int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                       TV* out, const msda_opts* opts) {
  if (!value || !offsets || !logits || !ref || !out) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "reference_points last dim must be 2 or 4, got %d", ref_dim);
  Problem pb;
  int rc = make_problem(s, shapes, start, batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, opts, &pb);
  if (rc != MSDA_OK) return rc;
  if (!fused_supported(pb, sizeof(TV)) || !aligned(value, 16) || !aligned(offsets, 16) || !aligned(logits, 16) || !aligned(out, 16))
    return fail(MSDA_ERR_UNSUPPORTED, "no fused-prologue kernel for this problem (needs head_dim 32, 4 points, 3-5 levels, the "
                                      "window or split kernel family and default flags)");
  pb.fz = MsdaFused{ref, ref_dim};
  return msda::fwd_d32<TV>(s, pb, value, offsets, logits, out);
}

template <typename TV>
int backward_fused_impl(cudaStream_t s, const TV* grad_out, const TV* value, const int64_t* shapes, const int64_t* start,
                        const float* offsets, const float* logits, const float* ref, int ref_dim, int batch,
                        int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                        float* gv, float* g_offsets, float* g_logits, const msda_opts* opts) {
  if (!grad_out || !value || !offsets || !logits || !ref || !gv || !g_offsets || !g_logits)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "reference_points last dim must be 2 or 4, got %d", ref_dim);
  Problem pb;
  int rc = make_problem(s, shapes, start, batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, opts, &pb);
  if (rc != MSDA_OK) return rc;
  if (!fused_supported(pb, sizeof(TV)) || !aligned(value, 16) || !aligned(gv, 16) || !aligned(offsets, 16) ||
      !aligned(logits, 16) || !aligned(grad_out, 16) || !aligned(g_offsets, 16) || !aligned(g_logits, 16))
    return fail(MSDA_ERR_UNSUPPORTED, "no fused-prologue kernel for this problem (needs head_dim 32, 4 points, 3-5 levels, the "
                                      "window or split kernel family and default flags)");
  pb.fz = MsdaFused{ref, ref_dim};
  if (!(pb.flags & MSDA_FLAG_GRAD_VALUE_PREZEROED)) {
    // the library's own fill kernel, which the backward kernel is launched behind programmatically: its front end /
    // gather phase overlaps the fill (see backward_impl)
    static const bool no_pdl = getenv("MSDA_B200_NO_PDL") && atoi(getenv("MSDA_B200_NO_PDL")) != 0;
    const size_t gv_bytes = (size_t)batch * spatial_size * num_heads * channels * sizeof(float);
    if (!no_pdl) {
      if ((rc = msda::zero_fill_pdl(s, gv, gv_bytes))) return rc;
      pb.pdl_after_fill = true;
    } else if ((rc = check_cuda(cudaMemsetAsync(gv, 0, gv_bytes, s), "zero-fill of grad_value"))) {
      return rc;
    }
  }
  return msda::kernel_family(pb) == MSDA_KERNEL_WINDOW
             ? msda::bwd_d32_win<TV>(s, pb, grad_out, value, offsets, logits, gv, g_offsets, g_logits)
             : msda::bwd_d32<TV, true>(s, pb, grad_out, value, offsets, logits, gv, g_offsets, g_logits);
}

}  // namespace

extern "C" {

int msda_forward_fused_f32(msda_stream_t stream, const float* value, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const float* sampling_offsets, const float* attn_logits,
                           const float* reference_points, int ref_dim, int batch, int spatial_size, int num_heads,
                           int channels, int num_levels, int num_query, int num_point, float* out, const msda_opts* opts) {
  return forward_fused_impl<float>((cudaStream_t)stream, value, spatial_shapes, level_start_index, sampling_offsets,
                                   attn_logits, reference_points, ref_dim, batch, spatial_size, num_heads, channels,
                                   num_levels, num_query, num_point, out, opts);
}
int msda_forward_fused_bf16(msda_stream_t stream, const uint16_t* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const float* sampling_offsets, const float* attn_logits,
                            const float* reference_points, int ref_dim, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point, uint16_t* out, const msda_opts* opts) {
  return forward_fused_impl<__nv_bfloat16>((cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(value), spatial_shapes,
                                           level_start_index, sampling_offsets, attn_logits, reference_points, ref_dim, batch,
                                           spatial_size, num_heads, channels, num_levels, num_query, num_point,
                                           reinterpret_cast<__nv_bfloat16*>(out), opts);
}
int msda_backward_fused_f32(msda_stream_t stream, const float* grad_out, const float* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const float* sampling_offsets, const float* attn_logits,
                            const float* reference_points, int ref_dim, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point, float* grad_value,
                            float* grad_sampling_offsets, float* grad_attn_logits, const msda_opts* opts) {
  return backward_fused_impl<float>((cudaStream_t)stream, grad_out, value, spatial_shapes, level_start_index, sampling_offsets,
                                    attn_logits, reference_points, ref_dim, batch, spatial_size, num_heads, channels, num_levels,
                                    num_query, num_point, grad_value, grad_sampling_offsets, grad_attn_logits, opts);
}
int msda_backward_fused_bf16(msda_stream_t stream, const uint16_t* grad_out, const uint16_t* value,
                             const int64_t* spatial_shapes, const int64_t* level_start_index, const float* sampling_offsets,
                             const float* attn_logits, const float* reference_points, int ref_dim, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                             float* grad_value, float* grad_sampling_offsets, float* grad_attn_logits, const msda_opts* opts) {
  return backward_fused_impl<__nv_bfloat16>((cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(grad_out),
                                            reinterpret_cast<const __nv_bfloat16*>(value), spatial_shapes, level_start_index,
                                            sampling_offsets, attn_logits, reference_points, ref_dim, batch, spatial_size,
                                            num_heads, channels, num_levels, num_query, num_point, grad_value,
                                            grad_sampling_offsets, grad_attn_logits, opts);
}

int msda_forward_f32(msda_stream_t stream, const float* value, const int64_t* spatial_shapes,
                     const int64_t* level_start_index, const float* sampling_loc,
                     const float* attn_weight, int batch, int spatial_size, int num_heads,
                     int channels, int num_levels, int num_query, int num_point, float* out,
                     const msda_opts* opts) {
  return forward_impl<float, float>((cudaStream_t)stream, value, spatial_shapes, level_start_index,
                             sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
                             num_levels, num_query, num_point, out, opts);
}

int msda_forward_f64(msda_stream_t stream, const double* value, const int64_t* spatial_shapes,
                     const int64_t* level_start_index, const double* sampling_loc,
                     const double* attn_weight, int batch, int spatial_size, int num_heads,
                     int channels, int num_levels, int num_query, int num_point, double* out,
                     const msda_opts* opts) {
  return forward_impl<double, double>((cudaStream_t)stream, value, spatial_shapes, level_start_index,
                              sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
                              num_levels, num_query, num_point, out, opts);
}

int msda_backward_f32(msda_stream_t stream, const float* grad_out, const float* value,
                      const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const float* sampling_loc, const float* attn_weight, int batch,
                      int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, float* grad_value, float* grad_sampling_loc,
                      float* grad_attn_weight, const msda_opts* opts) {
  return backward_impl<float, float>((cudaStream_t)stream, grad_out, value, spatial_shapes,
                              level_start_index, sampling_loc, attn_weight, batch, spatial_size,
                              num_heads, channels, num_levels, num_query, num_point, grad_value,
                              grad_sampling_loc, grad_attn_weight, opts);
}

int msda_backward_f64(msda_stream_t stream, const double* grad_out, const double* value,
                      const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const double* sampling_loc, const double* attn_weight, int batch,
                      int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, double* grad_value, double* grad_sampling_loc,
                      double* grad_attn_weight, const msda_opts* opts) {
  return backward_impl<double, double>((cudaStream_t)stream, grad_out, value, spatial_shapes,
                               level_start_index, sampling_loc, attn_weight, batch, spatial_size,
                               num_heads, channels, num_levels, num_query, num_point, grad_value,
                               grad_sampling_loc, grad_attn_weight, opts);
}

int msda_forward_bf16(msda_stream_t stream, const uint16_t* value, const int64_t* spatial_shapes,
                      const int64_t* level_start_index, const float* sampling_loc,
                      const float* attn_weight, int batch, int spatial_size, int num_heads,
                      int channels, int num_levels, int num_query, int num_point, uint16_t* out,
                      const msda_opts* opts) {
  return forward_impl<__nv_bfloat16, float>(
      (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(value), spatial_shapes,
      level_start_index, sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
      num_levels, num_query, num_point, reinterpret_cast<__nv_bfloat16*>(out), opts);
}

int msda_backward_bf16(msda_stream_t stream, const uint16_t* grad_out, const uint16_t* value,
                       const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const float* sampling_loc, const float* attn_weight, int batch,
                       int spatial_size, int num_heads, int channels, int num_levels,
                       int num_query, int num_point, float* grad_value, float* grad_sampling_loc,
                       float* grad_attn_weight, const msda_opts* opts) {
  return backward_impl<__nv_bfloat16, float>(
      (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(grad_out),
      reinterpret_cast<const __nv_bfloat16*>(value), spatial_shapes, level_start_index, sampling_loc,
      attn_weight, batch, spatial_size, num_heads, channels, num_levels, num_query, num_point,
      grad_value, grad_sampling_loc, grad_attn_weight, opts);
}

int msda_debug_corners_f32(msda_stream_t stream, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const float* sampling_loc, int batch,
                           int num_heads, int num_levels, int num_query, int num_point,
                           int32_t* corners, const msda_opts* opts) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!sampling_loc || !corners) return fail(MSDA_ERR_INVALID_ARGUMENT, "NULL tensor pointer");
  if (batch < 0 || num_heads < 1 || num_query < 0 || num_point < 1)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "bad dimensions");
  MsdaLevels lv;
  int rc = resolve_levels(s, spatial_shapes, level_start_index, num_levels, INT32_MAX, opts, &lv);
  if (rc != MSDA_OK) return rc;
  lv.coord_fma = (opt_flags(opts) & MSDA_FLAG_COORDS_FMA) ? 1 : 0;
  const long long n = (long long)batch * num_query * num_heads * num_levels * num_point;
  if (n == 0) return MSDA_OK;
  return msda::corners_probe(s, lv, sampling_loc, corners, n, num_levels, num_point);
}

size_t msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels,
                                     int num_levels, int num_query, int num_point) {
  return msda::det_workspace_bytes(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
}

int msda_has_fast_path(int dtype_bytes, int channels, int num_levels, int num_point) {
  return fast_shape(dtype_bytes, channels, num_levels, num_point) ? 1 : 0;
}

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

#define MSDA_STR_(x) #x
#define MSDA_STR(x) MSDA_STR_(x)
const char* msda_build_info(void) { return "msda_b200 abi " MSDA_STR(MSDA_ABI_VERSION) " sm_100a " __DATE__ " " __TIME__; }

const char* msda_last_error(void) { return g_err; }

uint64_t msda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
