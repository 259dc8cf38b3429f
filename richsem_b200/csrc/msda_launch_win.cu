// msda_launch_win.cu — instantiates and launches the shared-memory window kernels (msda_d32_win.cuh).
#include "msda_host.h"
#include "msda_d32_win.cuh"
#include "msda_d32_gv.cuh"

#include <atomic>

namespace msda {
namespace {

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute belongs to (function, device), so it is
// set once per device the kernel is launched on (one process may drive several GPUs), not once per process.
template <auto kKern>
int ensure_dynamic_smem(int bytes, const char* what) {
  static std::atomic<unsigned long long> done{0};  // one instance per kernel (the function is the template argument)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice");
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return MSDA_OK;
  e = cudaFuncSetAttribute(kKern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return check_cuda(e, what);
  done.fetch_or(bit, std::memory_order_release);
  return MSDA_OK;
}

template <typename VT, int kL, int kM>
int launch_fwd_win(cudaStream_t s, const Problem& pb, const VT* value, const float* loc,
                   const float* attw, VT* out) {
  using Cfg = WinCfg<VT, kL, kWinPoolFwd>;
  constexpr auto kern = msda_fwd_d32_win_kernel<VT, kL, kM>;
  if (int rc = ensure_dynamic_smem<kern>(Cfg::FWD_SMEM, "cudaFuncSetAttribute(msda_fwd_d32_win_kernel)")) return rc;
  const int tiles = (pb.order_len + kWinTileQ - 1) / kWinTileQ;
  dim3 grid(tiles * pb.d.num_heads, pb.d.batch);
  kern<<<grid, kWinThreads, Cfg::FWD_SMEM, s>>>(value, loc, attw, out, pb.order, pb.order_len, pb.lv,
                                                      pb.d.spatial_size, pb.d.num_heads, pb.d.num_query);
  return after_launch("msda_fwd_d32_win_kernel");
}


template <typename VT>
WinBwdArgs make_bwd_args(const Problem& pb, const VT* go, const VT* value, const float* loc, const float* attw,
                         float* gv, float* gl, float* ga) {
  WinBwdArgs a;
  a.grad_out = go; a.value = value; a.loc = loc; a.attw = attw;
  a.grad_value = gv; a.grad_loc = gl; a.grad_attw = ga;
  a.order = pb.order; a.order_len = pb.order_len;
  a.S = pb.d.spatial_size; a.M = pb.d.num_heads; a.Lq = pb.d.num_query;
  a.gv64 = nullptr; a.maxbits = nullptr;
  a.fz = pb.fz;
  return a;
}

template <typename VT, int kL, int kM, bool kDet, bool kFused = false>
int launch_bwd_win(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
                   const float* attw, float* gv, float* gl, float* ga, long long* gv64, const unsigned* maxbits) {
  using Cfg = WinCfg<VT, kL, kWinPoolBwd>;
  constexpr auto kern = msda_bwd_d32_win_kernel<VT, kL, kM, kDet, kFused>;
  constexpr int kSmem = (kDet || MSDA_WIN_MATCH_RANK) ? Cfg::BWD_DET_SMEM : Cfg::BWD_SMEM;
  if (int rc = ensure_dynamic_smem<kern>(kSmem, "cudaFuncSetAttribute(msda_bwd_d32_win_kernel)")) return rc;
  const int tiles = (pb.order_len + kWinTileQ - 1) / kWinTileQ;
  dim3 grid(tiles * pb.d.num_heads, pb.d.batch);
  WinBwdArgs a = make_bwd_args(pb, go, value, loc, attw, gv, gl, ga);
  a.gv64 = gv64; a.maxbits = maxbits;
  if (!kDet && pb.pdl_after_fill) {
    // programmatic dependent launch behind msda_zero_fill_kernel: the front end overlaps the fill
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWinThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, pb.lv);
    if (e != cudaSuccess) return check_cuda(e, "launch of msda_bwd_d32_win_kernel");
    return after_launch("msda_bwd_d32_win_kernel");
  }
  kern<<<grid, kWinThreads, kSmem, s>>>(a, pb.lv);
  return after_launch("msda_bwd_d32_win_kernel");
}

// persistent, warp-specialised variant: one 2 x kWinThreads block per SM
template <typename VT, int kL, int kM>
int launch_bwd_ws(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
                  const float* attw, float* gv, float* gl, float* ga) {
  using Cfg = WinCfg<VT, kL, kWinPoolBwd>;
  constexpr auto kern = msda_bwd_d32_ws_kernel<VT, kL, kM>;
  if (int rc = ensure_dynamic_smem<kern>(Cfg::BWD_WS_SMEM, "cudaFuncSetAttribute(msda_bwd_d32_ws_kernel)")) return rc;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaDeviceGetAttribute(MultiProcessorCount)");
  const int tiles = (pb.order_len + kWinTileQ - 1) / kWinTileQ;
  const long long total = (long long)tiles * pb.d.num_heads * pb.d.batch;
  const int grid = (int)(total < sms ? total : sms);
  kern<<<grid, 2 * kWinThreads, Cfg::BWD_WS_SMEM, s>>>(make_bwd_args(pb, go, value, loc, attw, gv, gl, ga), pb.lv, tiles,
                                                      pb.d.batch);
  return after_launch("msda_bwd_d32_ws_kernel");
}

template <typename VT, int kL, int kM>
int launch_gradvalue(cudaStream_t s, const Problem& pb, const VT* go, const float* loc, const float* attw, float* gv) {
  using Cfg = GvCfg<kL>;
  constexpr auto kern = msda_gradvalue_d32_kernel<VT, kL, kM>;
  if (int rc = ensure_dynamic_smem<kern>(Cfg::SMEM_BYTES, "cudaFuncSetAttribute(msda_gradvalue_d32_kernel)")) return rc;
  const int tiles = (pb.order_len + kWinTileQ - 1) / kWinTileQ;
  dim3 grid(tiles * pb.d.num_heads, pb.d.batch);
  kern<<<grid, kWinThreads, Cfg::SMEM_BYTES, s>>>(go, loc, attw, gv, pb.order, pb.order_len, pb.lv,
                                                  pb.d.spatial_size, pb.d.num_heads, pb.d.num_query);
  return after_launch("msda_gradvalue_d32_kernel");
}

#define MSDA_SWITCH_L(L_, CALL)                                                              \
  switch (L_) {                                                                              \
    case 1: return CALL(1);                                                                  \
    case 2: return CALL(2);                                                                  \
    case 3: return CALL(3);                                                                  \
    case 4: return CALL(4);                                                                  \
    case 5: return CALL(5);                                                                  \
    case 6: return CALL(6);                                                                  \
    default: return fail(MSDA_ERR_UNSUPPORTED, "no tuned kernel for num_levels=%d", L_);      \
  }

}  // namespace

template <typename VT>
int fwd_d32_win(cudaStream_t s, const Problem& pb, const VT* value, const float* loc, const float* attw, VT* out) {
  // the DINO / RichSem configuration (8 heads, 4 levels) gets the head count baked in
  if (pb.d.num_heads == 8 && pb.d.num_levels == 4) return launch_fwd_win<VT, 4, 8>(s, pb, value, loc, attw, out);
#define CALL(L) launch_fwd_win<VT, L, 0>(s, pb, value, loc, attw, out)
  MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
}

template <typename VT>
int bwd_d32_win(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
                const float* attw, float* gv, float* gl, float* ga) {
  if (pb.flags & MSDA_FLAG_BWD_WS) {
    if (pb.d.num_heads == 8 && pb.d.num_levels == 4) return launch_bwd_ws<VT, 4, 8>(s, pb, go, value, loc, attw, gv, gl, ga);
#define CALL(L) launch_bwd_ws<VT, L, 0>(s, pb, go, value, loc, attw, gv, gl, ga)
    MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
  }
  if (pb.fz.ref_dim) {  // fused prologue
    if (pb.d.num_heads == 8 && pb.d.num_levels == 4)
      return launch_bwd_win<VT, 4, 8, false, true>(s, pb, go, value, loc, attw, gv, gl, ga, nullptr, nullptr);
#define CALL(L) launch_bwd_win<VT, L, 0, false, true>(s, pb, go, value, loc, attw, gv, gl, ga, nullptr, nullptr)
    MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
  }
  if (pb.d.num_heads == 8 && pb.d.num_levels == 4)
    return launch_bwd_win<VT, 4, 8, false>(s, pb, go, value, loc, attw, gv, gl, ga, nullptr, nullptr);
#define CALL(L) launch_bwd_win<VT, L, 0, false>(s, pb, go, value, loc, attw, gv, gl, ga, nullptr, nullptr)
  MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
}

// Deterministic backward for large problems.  Workspace: [256 B header: max|grad_out| bits, max|attn_weight|
// bits][int64 fixed-point accumulators, one per grad_value element].
template <typename VT>
int bwd_d32_win_det(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
                    const float* attw, float* gv, float* gl, float* ga, void* ws, size_t ws_bytes) {
  const size_t n_gv = (size_t)pb.d.batch * pb.d.spatial_size * pb.d.num_heads * 32;
  const size_t need = 256 + n_gv * 8;
  if (!ws || ws_bytes < need)
    return fail(MSDA_ERR_WORKSPACE, "deterministic mode needs %zu workspace bytes, got %zu", need, ws_bytes);
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(MSDA_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  unsigned* maxbits = static_cast<unsigned*>(ws);
  long long* gv64 = reinterpret_cast<long long*>(static_cast<char*>(ws) + 256);
  int rc = check_cuda(cudaMemsetAsync(ws, 0, need, s), "workspace clear");
  if (rc) return rc;
  const size_t n_go = (size_t)pb.d.batch * pb.d.num_query * pb.d.num_heads * 32;
  const size_t n_aw = (size_t)pb.d.batch * pb.d.num_query * pb.d.num_heads * pb.d.num_levels * pb.d.num_point;
  msda_maxabs_kernel<VT><<<148 * 8, 256, 0, s>>>(go, n_go, attw, n_aw, maxbits);
  if ((rc = after_launch("msda_maxabs_kernel"))) return rc;
  if (pb.d.num_heads == 8 && pb.d.num_levels == 4) {
    rc = launch_bwd_win<VT, 4, 8, true>(s, pb, go, value, loc, attw, gv, gl, ga, gv64, maxbits);
  } else {
    rc = [&]() -> int {
#define CALL(L) launch_bwd_win<VT, L, 0, true>(s, pb, go, value, loc, attw, gv, gl, ga, gv64, maxbits)
      MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
    }();
  }
  if (rc) return rc;
  msda_fixed_to_float_kernel<<<148 * 8, 256, 0, s>>>(gv64, gv, n_gv, maxbits, pb.d.num_query, pb.d.num_levels * pb.d.num_point);
  return after_launch("msda_fixed_to_float_kernel");
}

template <typename VT>
int gradvalue_d32(cudaStream_t s, const Problem& pb, const VT* go, const float* loc, const float* attw, float* gv) {
  if (pb.d.num_heads == 8 && pb.d.num_levels == 4) return launch_gradvalue<VT, 4, 8>(s, pb, go, loc, attw, gv);
#define CALL(L) launch_gradvalue<VT, L, 0>(s, pb, go, loc, attw, gv)
  MSDA_SWITCH_L(pb.d.num_levels, CALL)
#undef CALL
}
template int gradvalue_d32<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, float*);
template int gradvalue_d32<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const float*, const float*,
                                          float*);

template int bwd_d32_win_det<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, const float*,
                                    float*, float*, float*, void*, size_t);
template int bwd_d32_win_det<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const __nv_bfloat16*,
                                            const float*, const float*, float*, float*, float*, void*, size_t);
template int fwd_d32_win<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, float*);
template int fwd_d32_win<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const float*,
                                        const float*, __nv_bfloat16*);

template int bwd_d32_win<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, const float*,
                                float*, float*, float*);
template int bwd_d32_win<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const __nv_bfloat16*,
                                        const float*, const float*, float*, float*, float*);

#ifdef MSDA_WIN_TIMING
// debug builds only: copies and clears the phase-timing accumulators
extern "C" int msda_debug_win_timing(unsigned long long* out16) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_win_timing, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_win_timing, z, sizeof(z));
  return (int)e;
}
#endif

}  // namespace msda
