#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long addrz2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__global__ void k(const float2* __restrict__ in, const float* __restrict__ w, float2* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float2 p00 = in[i], p10 = in[i + 1], p01 = in[i + 640], p11 = in[i + 641];
  float fu = w[i], fv = w[i + 7];
  float gu = 1.0f - fu, gv = 1.0f - fv;
  unsigned long long a = *reinterpret_cast<unsigned long long*>(&p00), b = *reinterpret_cast<unsigned long long*>(&p10);
  unsigned long long c = *reinterpret_cast<unsigned long long*>(&p01), d = *reinterpret_cast<unsigned long long*>(&p11);
  unsigned long long r = fma2(pk(fv, fv), fma2(pk(fu, fu), d, mul2(pk(gu, gu), c)), mul2(pk(gv, gv), fma2(pk(fu, fu), b, mul2(pk(gu, gu), a))));
  r = addrz2(r, pk(8388608.f, 8388608.f));
  out[i] = *reinterpret_cast<float2*>(&r);
}
