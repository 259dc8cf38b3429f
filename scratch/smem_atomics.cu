// microbenchmark: shared-memory accumulation primitives (throughput per SM), sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 512            // 512 rows x 32 floats = 64 KB tile
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, unsigned seed) {
  extern __shared__ float tile[];
  for (int i = threadIdx.x; i < ROWS * 32 * (MODE == 3 ? 2 : 1); i += 256) tile[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned s = seed + warp * 7919u + blockIdx.x * 104729u;
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const int row = (s >> 10) % ROWS;          // warp-uniform random row, 32 lanes = 32 channels
    const float v = 1.0f + lane;
    if (MODE == 0) {            // plain read-modify-write (racy; upper bound)
      tile[row * 32 + lane] += v;
    } else if (MODE == 1) {     // float atomicAdd (CAS loop on sm_100a)
      atomicAdd(&tile[row * 32 + lane], v);
    } else if (MODE == 2) {     // int32 fixed-point atomicAdd (native ATOMS.ADD)
      atomicAdd(reinterpret_cast<int*>(tile) + row * 32 + lane, (int)(v * 1024.f));
    } else if (MODE == 3) {     // int64 fixed-point atomicAdd
      atomicAdd(reinterpret_cast<unsigned long long*>(tile) + row * 32 + lane, (unsigned long long)(v * 1024.f));
    }
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < ROWS * 32; i += 256) acc += tile[i];
  if (acc == 123.456f) out[0] = acc;
}
__global__ void __launch_bounds__(256) kred(float* g, int iters, unsigned seed, int rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned s = seed + warp * 7919u + blockIdx.x * 104729u;
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const int grp = lane >> 3;
    const unsigned row = ((s >> 8) + grp * 977u) % rows;
    float* p = g + (size_t)row * 32 + (lane & 7) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
  }
}
template <int MODE> float run(int iters) {
  float* out; cudaMalloc(&out, 4);
  size_t sm = ROWS * 32 * 4 * (MODE == 3 ? 2 : 1);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, 256, sm>>>(out, 10, 1); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148, 256, sm>>>(out, iters, 1); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); cudaFree(out); return ms;
}
int main() {
  int iters = 20000; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const char* names[] = {"plain RMW (LDS+FADD+STS)", "float atomicAdd (CAS)", "int32 atomicAdd", "int64 atomicAdd"};
  float ms[4] = {run<0>(iters), run<1>(iters), run<2>(iters), run<3>(iters)};
  for (int m = 0; m < 4; ++m) {
    double row_ops_per_sm = 8.0 * iters;   // 8 warps x iters rows of 128 B
    double cyc = ms[m] * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.2f cycles per 128-B row-add per SM (clock %d kHz)\n", names[m], ms[m], cyc / row_ops_per_sm, clk);
  }
  // global RED baseline: 148*4 blocks, rows spread over 355k rows (45 MB) -> L2
  float* g; int rows = 355568; cudaMalloc(&g, (size_t)rows * 128); cudaMemset(g, 0, (size_t)rows * 128);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  kred<<<148 * 4, 256>>>(g, 100, 3, rows); cudaDeviceSynchronize();
  cudaEventRecord(a); kred<<<148 * 4, 256>>>(g, 5000, 3, rows); cudaEventRecord(b); cudaEventSynchronize(b);
  float msr; cudaEventElapsedTime(&msr, a, b);
  double rows_total = 148.0 * 4 * 8 * 5000 * 4;
  printf("global RED.v4 random rows    %8.3f ms  %6.2f cycles per 128-B row-add per SM, %.2f TB/s payload\n", msr,
         msr * 1e-3 * clk * 1e3 / (rows_total / 148), rows_total * 128 / (msr * 1e-3) / 1e12);
  return 0;
}
