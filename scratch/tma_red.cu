// microbenchmark: cp.reduce.async.bulk (TMA reduce-add, smem -> global) for scattered 128-B rows vs RED.v4
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ROW_BYTES, int DEPTH>
__global__ void __launch_bounds__(256) ktma(float* g, int iters, unsigned seed, int rows) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int ROWS_PER_STEP = 512 / ROW_BYTES * 1;  // one warp-wide STS.128 = 512 B
  unsigned char* wbase = sm + warp * DEPTH * 512;
  unsigned s = seed + warp * 7919u + blockIdx.x * 104729u;
  for (int it = 0; it < iters; ++it) {
    unsigned char* slot = wbase + (it % DEPTH) * 512;
    if (it >= DEPTH) {  // slot reuse: wait until the bulk op that read it is done reading
      if (lane < ROWS_PER_STEP) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
      __syncwarp();
    }
    reinterpret_cast<float4*>(slot)[lane] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    s = s * 1664525u + 1013904223u;
    if (lane < ROWS_PER_STEP) {
      const unsigned row = ((s >> 8) + lane * 977u) % rows;
      float* dst = g + (size_t)row * (ROW_BYTES / 4);
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(slot + lane * ROW_BYTES);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(ROW_BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane < ROWS_PER_STEP) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
template <int RB, int DEPTH> void run(float* g, int rows128, int blocks_per_sm, int iters) {
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int smem = 8 * DEPTH * 512;
  cudaFuncSetAttribute(ktma<RB, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int rows = rows128 * 128 / RB;
  ktma<RB, DEPTH><<<148 * blocks_per_sm, 256, smem>>>(g, 50, 3, rows); 
  cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return; }
  cudaEventRecord(a); ktma<RB, DEPTH><<<148 * blocks_per_sm, 256, smem>>>(g, iters, 3, rows); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double bytes = 148.0 * blocks_per_sm * 8 * iters * 512.0;
  printf("TMA reduce rows of %4d B depth %d, %d blocks/SM: %8.3f ms  %.2f TB/s payload, %.2f cycles per 128 B per SM\n", RB, DEPTH, blocks_per_sm, ms,
         bytes / (ms * 1e-3) / 1e12, ms * 1e-3 * clk * 1e3 / (bytes / 128 / 148));
}
int main() {
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* g; int rows = 355568; cudaMalloc(&g, (size_t)rows * 128); cudaMemset(g, 0, (size_t)rows * 128);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int bps = 1; bps <= 8; bps *= 2) {
    kred<<<148 * bps, 256>>>(g, 100, 3, rows); cudaDeviceSynchronize();
    int iters = 20000 / bps;
    cudaEventRecord(a); kred<<<148 * bps, 256>>>(g, iters, 3, rows); cudaEventRecord(b); cudaEventSynchronize(b);
    float msr; cudaEventElapsedTime(&msr, a, b);
    double rows_total = 148.0 * bps * 8 * iters * 4;
    printf("RED.v4 random 128-B rows, %d blocks/SM: %8.3f ms  %.2f TB/s payload, %.2f cycles per row per SM\n", bps, msr,
           rows_total * 128 / (msr * 1e-3) / 1e12, msr * 1e-3 * clk * 1e3 / (rows_total / 148));
  }
  run<128, 4>(g, rows, 1, 5000); run<128, 4>(g, rows, 4, 2000); run<128, 8>(g, rows, 4, 2000);
  run<256, 4>(g, rows, 4, 2000); run<512, 4>(g, rows, 4, 2000);
  float h[8]; cudaMemcpy(h, g, 32, cudaMemcpyDeviceToHost); printf("check %f %f\n", h[0], h[1]);
  return 0;
}
