// microbenchmark: staging a bh x bw window of 128-byte value rows (row stride 1 KB) into shared memory
//   mode 0: cp.async.cg 16 B per lane (LDGSTS)
//   mode 1: cp.async.bulk 1-D, 128 B per row, one op per thread, mbarrier complete_tx
//   mode 2: cp.async.bulk.tensor 4-D box {32, 1, bw, bh}, one op per block
//   gather>0: every thread additionally gathers `gather` rows (LDS.128) from a second resident window per
//   iteration, to see whether staging and gathering overlap
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
constexpr int H = 100, W = 167, M = 8, BH = 16, BW = 16, ROWS = BH * BW;
constexpr int BXW = 4;  // columns per small TMA box (mode 3)
__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT;\n DONE:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ value, const __grid_constant__ CUtensorMap tmap,
                                         const __grid_constant__ CUtensorMap tmap_small, float* out, int iters, int gather) {
  extern __shared__ __align__(128) unsigned char smraw[];
  float4* win = reinterpret_cast<float4*>(smraw);                   // staged window
  float4* res = reinterpret_cast<float4*>(smraw + ROWS * 128);      // resident window for the gather
  __shared__ __align__(8) unsigned long long bar;
  const int t = threadIdx.x, lane = t & 31, j = lane & 7;
  for (int i = t; i < ROWS * 8; i += 256) res[i] = make_float4(1, 2, 3, 4);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned sb = 4242u + blockIdx.x * 104729u;
  unsigned sg = 12345u + (t >> 3) * 7919u + blockIdx.x * 104729u;
  float4 acc = make_float4(0, 0, 0, 0);
  unsigned parity = 0;
  for (int it = 0; it < iters; ++it) {
    const int h0 = lcg(sb) % (H - BH), w0 = lcg(sb) % (W - BW), m = lcg(sb) % M, b = lcg(sb) & 1;
    const float* base = value + (size_t)b * H * W * M * 32 + m * 32;
    if (MODE == 0) {
      for (int i = t; i < ROWS * 8; i += 256) {
        const int r = i >> 3, rh = r / BW, rw = r % BW;
        const float* src = base + ((size_t)(h0 + rh) * W + w0 + rw) * (M * 32) + (i & 7) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(win + i)), "l"(src));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else if (MODE == 1) {
      if (t == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ROWS * 128) : "memory");
      __syncthreads();
      for (int r = t; r < ROWS; r += 256) {
        const int rh = r / BW, rw = r % BW;
        const float* src = base + ((size_t)(h0 + rh) * W + w0 + rw) * (M * 32);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(smem_u32(win + r * 8)),
                     "l"(src), "r"(smem_u32(&bar))
                     : "memory");
      }
    } else if (MODE == 3) {
      if (t == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ROWS * 128) : "memory");
      __syncthreads();
      constexpr int CH = BW / BXW;
      if (t < BH * CH) {
        const int rh = t / CH, ch = t % CH;
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                smem_u32(win + (rh * BW + ch * BXW) * 8)),
            "l"(&tmap_small), "r"(0), "r"(m), "r"(w0 + ch * BXW), "r"(h0 + rh), "r"(b), "r"(smem_u32(&bar))
            : "memory");
      }
    } else {
      if (t == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ROWS * 128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                smem_u32(win)),
            "l"(&tmap), "r"(0), "r"(m), "r"(w0), "r"(h0), "r"(b), "r"(smem_u32(&bar))
            : "memory");
      }
    }
    for (int u = 0; u < gather; ++u) {
      const int row = lcg(sg) % ROWS;
      const float4 v = res[row * 8 + j];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (MODE == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else { mbar_wait(smem_u32(&bar), parity); parity ^= 1; }
    __syncthreads();
    const float4 v = win[(it * 37 + t) % (ROWS * 8)];
    acc.x += v.x;
    __syncthreads();
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int MODE> void run(const char* name, const float* value, const CUtensorMap& tm, const CUtensorMap& tms, int bps, int iters, int gather) {
  float* out; cudaMalloc(&out, 16);
  size_t sm = 2 * ROWS * 128;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148 * bps, 256, sm>>>(value, tm, tms, out, 10, gather); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148 * bps, 256, sm>>>(value, tm, tms, out, iters, gather); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3 / (iters * bps);   // cycles per window per SM-slot
  printf("%-28s blocks/SM %d gather %3d rows/thread: %8.3f ms  %7.1f cycles per 256-row window per SM (%.2f / staged row; gather alone would be %.0f)  %s\n",
         name, bps, gather, ms, cyc, cyc / ROWS, gather * 32.0 * 1.04, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  float* value; size_t n = (size_t)2 * H * W * M * 32; cudaMalloc(&value, n * 4); cudaMemset(value, 0, n * 4);
  EncodeFn enc; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  CUtensorMap tm;
  cuuint64_t dims[5] = {32, M, W, H, 2};
  cuuint64_t strides[4] = {128, 1024, (cuuint64_t)W * 1024, (cuuint64_t)H * W * 1024};
  cuuint32_t box[5] = {32, 1, BW, BH, 1}, es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, value, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUtensorMap tms;
  cuuint32_t box_s[5] = {32, 1, BXW, 1, 1};
  CUresult r2 = enc(&tms, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, value, dims, strides, box_s, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d %d\n", (int)r, (int)r2);
  for (int bps = 1; bps <= 2; ++bps)
    for (int gather : {0, 64, 128}) {
      run<0>("cp.async 16 B", value, tm, tms, bps, 2000, gather);
      run<2>("tensor 5-D box 16x16", value, tm, tms, bps, 2000, gather);
      run<3>("tensor 5-D boxes 1x4 (64 ops)", value, tm, tms, bps, 2000, gather);
    }
  return 0;
}
