// micro-benchmark: issue/throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[16]; u64 p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) { float2 t = make_float2(a[2*i], a[2*i+1]); p[i] = *reinterpret_cast<u64*>(&t); }
  float2 ss = make_float2(s, s); u64 s2 = *reinterpret_cast<u64*>(&ss);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(a[i]), "f"(s), "f"(s));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], s2, s2);
    } else {  // mixed: 8 FFMA2 + 8 integer ops to see co-issue
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], s2, s2);
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(*(unsigned*)&a[i]) : "r"(*(unsigned*)&a[i]), "r"(it), "r"(i));
    }
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); r += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.0001f);
      else if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, 1.0001f);
      else k<2><<<148, warps * 32>>>(out, iters, 1.0001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_per_thread = (double)iters * 16;  // scalar-equivalent FMAs
    double tf = fma_per_thread * 148 * warps * 32 * 2 / (ms * 1e-3) / 1e12;
    printf("mode %d warps/SM %2d: %.3f ms  %.1f TFLOP/s fp32\n", mode, warps, ms, tf);
  }
  return 0;
}
