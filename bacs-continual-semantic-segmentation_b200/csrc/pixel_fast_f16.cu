// Instantiation of the register-resident per-pixel kernels for one storage type
// (one translation unit per type so that nvcc compiles them in parallel).
#include "pixel_fast.cuh"
#include "pixel_wce.cuh"

namespace bacs {
int launch_pixel_fast_f16(const PixelParams& p, const PixelPlan& plan, cudaStream_t s) {
  return launch_fast_dtype<__half>(p, plan, s);
}
int launch_pixel_wce_f16(const PixelParams& p, const PixelPlan& plan, cudaStream_t s) {
  return launch_wce_dtype<__half>(p, plan, s);
}
}  // namespace bacs
