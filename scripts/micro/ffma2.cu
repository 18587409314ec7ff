// microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], s, 0.5f + i);
    }
  } else {
    unsigned long long p[8], sc = pk(s, s);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], sc, pk(0.5f + 2 * i, 1.5f + 2 * i));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) upk(p[i], a[2 * i], a[2 * i + 1]);
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, 20000, 0.999f); else k<1><<<148 * 8, 256>>>(d, 20000, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fmas = 148.0 * 8 * 256 * 16 * 20000;
      printf("mode %d: %.3f ms  %.2f TFMA/s (%.1f TFLOP/s)  err=%s\n", mode, ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
