// micro-benchmark: legacy mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16), independent accumulators
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = threadIdx.x * 11, b1 = threadIdx.x * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float r = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) r += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode) for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, warps * 32>>>(out, iters); else k<1><<<148, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop_per_mma = mode == 0 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 16;
    const double tf = flop_per_mma * 8 * iters * warps * 148 / (ms * 1e-3) / 1e12;
    printf("%s warps/SM %2d: %.3f ms  %.0f TFLOP/s  (%.2f mma/clk/SM)\n", mode == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", warps, ms, tf,
           8.0 * iters * warps / (ms * 1e-3 * 1.965e9));
  }
  return 0;
}
