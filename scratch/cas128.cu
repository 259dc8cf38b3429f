// microbenchmark: shared-memory float4 accumulate with a 128-bit CAS loop vs scalar float atomicAdd
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 448
__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ void add_f4_cas128(float4* p, float4 v) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  float4 old = *p;
  while (true) {
    float4 nw = make_float4(old.x + v.x, old.y + v.y, old.z + v.z, old.w + v.w);
    float4 got;
    unsigned long long g01, g23;
    asm volatile(
        "{\n .reg .b128 c, n, d;\n"
        " mov.b128 c, {%2, %3};\n mov.b128 n, {%4, %5};\n"
        " atom.relaxed.cta.shared.cas.b128 d, [%6], c, n;\n"
        " mov.b128 {%0, %1}, d;\n}\n"
        : "=l"(g01), "=l"(g23)
        : "l"(*reinterpret_cast<unsigned long long*>(&old.x)), "l"(*reinterpret_cast<unsigned long long*>(&old.z)),
          "l"(*reinterpret_cast<unsigned long long*>(&nw.x)), "l"(*reinterpret_cast<unsigned long long*>(&nw.z)), "r"(a)
        : "memory");
    *reinterpret_cast<unsigned long long*>(&got.x) = g01;
    *reinterpret_cast<unsigned long long*>(&got.z) = g23;
    if (__float_as_uint(got.x) == __float_as_uint(old.x) && __float_as_uint(got.y) == __float_as_uint(old.y) &&
        __float_as_uint(got.z) == __float_as_uint(old.z) && __float_as_uint(got.w) == __float_as_uint(old.w)) break;
    old = got;
  }
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  extern __shared__ float4 win[];
  const int t = threadIdx.x, lane = t & 31, j = lane & 7;
  for (int i = t; i < ROWS * 8; i += 256) win[i] = make_float4(0, 0, 0, 0);
  __syncthreads();
  unsigned s = 12345u + (t >> 3) * 7919u + blockIdx.x * 104729u;
  for (int it = 0; it < iters; ++it) {
    const int row = lcg(s) % ROWS;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    if (MODE == 0) add_f4_cas128(&win[row * 8 + j], v);
    else if (MODE == 1) {
      float* p = reinterpret_cast<float*>(&win[row * 8 + j]);
      atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
    } else {
      float4 o = win[row * 8 + j]; o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w; win[row * 8 + j] = o;
    }
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = t; i < ROWS * 8; i += 256) acc += win[i].x + win[i].w;
  atomicAdd(out + MODE, acc);
}
template <int MODE> void run(const char* name, int bps, int iters) {
  float* out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
  size_t sm = ROWS * 128;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148 * bps, 256, sm>>>(out, 10); cudaDeviceSynchronize(); cudaMemset(out, 0, 16);
  cudaEventRecord(a); k<MODE><<<148 * bps, 256, sm>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float h[4]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  double expect = 148.0 * bps * 256.0 / 8 * iters * 8 * 5.0;  // groups*iters rows, each row: 8 lanes * (1+4)
  printf("%-36s blocks/SM %d %8.3f ms  %6.2f cycles per 128-B row-add per SM  sum=%.6g expect=%.6g err=%s\n", name, bps, ms,
         ms * 1e-3 * clk * 1e3 / (32.0 * iters * bps), h[MODE], expect, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  for (int bps = 1; bps <= 2; ++bps) {
    run<0>("float4 CAS.b128 loop", bps, 20000);
    run<1>("4x scalar float atomicAdd", bps, 20000);
    run<2>("plain float4 RMW (racy)", bps, 20000);
  }
}
