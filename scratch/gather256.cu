// microbenchmark: L1-resident row gather with 256-bit loads (sm_100: ld.global.v8.f32) vs 128-bit
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 448
__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ g, float* out, int iters) {
  const int t = threadIdx.x, lane = t & 31;
  const float* gb = g + (size_t)blockIdx.x * ROWS * 32;
  float acc = 0.f;
  if (MODE == 0) {  // 128-bit, 8 lanes per row, 4 rows per instruction
    unsigned s = 12345u + (t >> 3) * 7919u + blockIdx.x * 104729u;
    const int j = lane & 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = lcg(s) % ROWS;
        const float4 v = __ldg(reinterpret_cast<const float4*>(gb + row * 32) + j);
        acc += v.x + v.y + v.z + v.w;
      }
    }
  } else {  // 256-bit, 4 lanes per row, 8 rows per instruction
    unsigned s = 12345u + (t >> 2) * 7919u + blockIdx.x * 104729u;
    const int j = lane & 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = lcg(s) % ROWS;
        float v0, v1, v2, v3, v4, v5, v6, v7;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(v4), "=f"(v5), "=f"(v6), "=f"(v7)
                     : "l"(gb + row * 32 + j * 8));
        acc += v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int MODE> void run(const char* name, int bps, int iters, double rows_per_iter_per_block) {
  float* g; float* out;
  const int blocks = 148 * bps;
  cudaMalloc(&g, (size_t)blocks * ROWS * 128); cudaMemset(g, 0, (size_t)blocks * ROWS * 128);
  cudaMalloc(&out, 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, 256>>>(g, out, 10); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<blocks, 256>>>(g, out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-40s blocks/SM %d  %8.3f ms  %6.3f cycles per row per SM  %s\n", name, bps, ms,
         ms * 1e-3 * clk * 1e3 / (rows_per_iter_per_block * iters * bps), cudaGetErrorString(cudaGetLastError()));
  cudaFree(g); cudaFree(out);
}
int main() {
  for (int bps = 1; bps <= 4; ++bps) {
    run<0>("LDG.128 gather (4 rows / instr)", bps, 4000, 32.0 * 8);
    run<1>("LDG.256 gather (8 rows / instr)", bps, 4000, 64.0 * 8);
  }
}
