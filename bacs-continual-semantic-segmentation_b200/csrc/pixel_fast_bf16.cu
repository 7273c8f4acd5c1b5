// Instantiation of the register-resident per-pixel kernels for one storage type
// (one translation unit per type so that nvcc compiles them in parallel).
#include "pixel_fast.cuh"
#include "pixel_wce.cuh"

namespace bacs {
int launch_pixel_fast_bf16(const PixelParams& p, const PixelPlan& plan, cudaStream_t s) {
  return launch_fast_dtype<__nv_bfloat16>(p, plan, s);
}
int launch_pixel_wce_bf16(const PixelParams& p, const PixelPlan& plan, cudaStream_t s) {
  return launch_wce_dtype<__nv_bfloat16>(p, plan, s);
}
}  // namespace bacs
