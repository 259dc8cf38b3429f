// msda_launch_win.cu — instantiates and launches the shared-memory window backward (msda_d32_win.cuh).
#include "msda_host.h"
#include "msda_d32_win.cuh"

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
  using Cfg = WinCfg<VT, kL>;
  constexpr auto kern = msda_bwd_d32_win_kernel<VT, kL, kM, kDet, kFused>;
#ifdef MSDA_WIN_EXTRA_SMEM  // occupancy experiments: pad the block's shared memory (e.g. 30000 -> one block per SM)
  constexpr int kSmem = (kDet ? Cfg::BWD_DET_SMEM : Cfg::BWD_SMEM) + MSDA_WIN_EXTRA_SMEM;
#else
  constexpr int kSmem = kDet ? Cfg::BWD_DET_SMEM : Cfg::BWD_SMEM;
#endif
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

#define MSDA_SWITCH_L(L_, CALL)                                                              \
  switch (L_) {                                                                              \
    case 3: return CALL(3);                                                                  \
    case 4: return CALL(4);                                                                  \
    case 5: return CALL(5);                                                                  \
    default: return fail(MSDA_ERR_UNSUPPORTED, "no tuned kernel for num_levels=%d", L_);      \
  }

}  // namespace

template <typename VT>
int bwd_d32_win(cudaStream_t s, const Problem& pb, const VT* go, const VT* value, const float* loc,
                const float* attw, float* gv, float* gl, float* ga) {
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

template int bwd_d32_win_det<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, const float*,
                                    float*, float*, float*, void*, size_t);
template int bwd_d32_win_det<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const __nv_bfloat16*,
                                            const float*, const float*, float*, float*, float*, void*, size_t);
template int bwd_d32_win<float>(cudaStream_t, const Problem&, const float*, const float*, const float*, const float*,
                                float*, float*, float*);
template int bwd_d32_win<__nv_bfloat16>(cudaStream_t, const Problem&, const __nv_bfloat16*, const __nv_bfloat16*,
                                        const float*, const float*, float*, float*, float*);

#ifdef MSDA_WIN_TIMING
// debug builds only: copies and clears the phase-timing accumulators
extern "C" int msda_debug_win_timing(unsigned long long* out32) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out32, g_win_timing, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {0};
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_win_timing, z, sizeof(z));
  return (int)e;
}
#endif

}  // namespace msda
