// Pixel x class-prototype squared distance on the 5th-generation tensor cores (north_star: "tensor cores (bf16 UMMA) only
// for the pixel-by-prototype distance when K*C is large enough (ADE20K's 150 classes)").
//
//   dist2[b, k, p] = |f_bp|^2 + |c_k|^2 - 2 <f_bp, c_k>      f = features [B, D, h*w] (NCHW), c = class prototypes [Kc, D]
//
// The reference arithmetic has no such product (its per-task heads are a weighted L1 of sigmoids, networks/
// bg_detector.py:17-40; SDR's per-class terms, loss/sdr.py:120-200, touch each pixel's own class only), so this is
// the OPTIONAL per-class prototype family entry, not part of the BACS step.  The dot products are one batched GEMM
// per call: M = h*w pixels of an image (contiguous in NCHW -> an M-major A operand, no transpose), N = classes,
// K = D, batch = images, bf16 operands / fp32 accumulation in tensor memory: tcgen05.mma (SASS UTCHMMA) fed by TMA
// (UTMALDG), accumulators read back with tcgen05.ld (LDTM), instantiated from the CUTLASS 4 SM100 collectives
// (vendored header tree) inside this translation unit.  Two small kernels of ours add the norms and the arg-min.
#include <algorithm>

#include "common.cuh"

#ifdef BACS_HAVE_CUTLASS
#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

namespace {
using namespace cute;
using ElementA = cutlass::bfloat16_t;
using ElementB = cutlass::bfloat16_t;
using ElementC = float;
using ElementAcc = float;
using LayoutA = cutlass::layout::ColumnMajor;  // pixels contiguous (NCHW features)
using LayoutB = cutlass::layout::ColumnMajor;  // D contiguous (prototype rows)
using LayoutC = cutlass::layout::ColumnMajor;  // pixels contiguous: [B, classes, h, w]
using ArchTag = cutlass::arch::Sm100;
using OpClass = cutlass::arch::OpClassTensorOp;
using TileShape = Shape<_128, _256, _64>;      // 128 pixels x up to 256 classes per CTA, one UMMA tile in N
using ClusterShape = Shape<_1, _1, _1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    ArchTag, OpClass, TileShape, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc,
    ElementC, LayoutC, 4, ElementC, LayoutC, 4, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    ArchTag, OpClass, ElementA, LayoutA, 8, ElementB, LayoutB, 8, ElementAcc, TileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
using GemmKernel =
    cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

typename Gemm::Arguments make_args(const void* A, const void* B, float* D, int M, int N, int K, int L) {
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  const StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, L));
  const StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, 1));  // shared by the batch
  const StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, L));
  const StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, cute::make_shape(M, N, L));
  return typename Gemm::Arguments{cutlass::gemm::GemmUniversalMode::kGemm,
                                  {M, N, K, L},
                                  {reinterpret_cast<const ElementA*>(A), sa, reinterpret_cast<const ElementB*>(B), sb},
                                  {{1.0f, 0.0f}, D, sc, D, sd}};
}
}  // namespace
#endif  // BACS_HAVE_CUTLASS

namespace bacs {

// cc[k] = |c_k|^2 (one warp per class)
__global__ void __launch_bounds__(256) class_sqnorm_kernel(const __nv_bfloat16* __restrict__ protos, int Kc, int D,
                                                           float* __restrict__ cc) {
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= Kc) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = __bfloat162float(protos[(int64_t)k * D + d]);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) cc[k] = s;
}

// A block takes 64 adjacent pixels of one image with 8 thread slices: slice s sums the squares of channels s, s+8, ...
// (every load is a 128-byte row segment), the slices meet in shared memory, then slice s turns the dot products of
// classes s, s+8, ... into distances in place and the slices' nearest classes are merged (ties -> lowest class).
constexpr int kCdPx = 64, kCdSlices = 8;
__global__ void __launch_bounds__(kCdPx* kCdSlices) class_distance_finish_kernel(const __nv_bfloat16* __restrict__ feat,
                                                                              int D, int hw, int Kc,
                                                                              const float* __restrict__ cc,
                                                                              float* __restrict__ dist2,
                                                                              int64_t* __restrict__ nearest) {
  __shared__ float s_part[kCdSlices][kCdPx];
  __shared__ int s_arg[kCdSlices][kCdPx];
  const int b = blockIdx.y;
  const int px = threadIdx.x % kCdPx, sl = threadIdx.x / kCdPx;
  const int p = blockIdx.x * kCdPx + px;
  const bool live = p < hw;
  float acc = 0.f;
  if (live) {
    const __nv_bfloat16* f = feat + (int64_t)b * D * hw + p;
    float xa = 0.f, xb = 0.f;
    int d = sl;
    for (; d + kCdSlices < D; d += 2 * kCdSlices) {
      const float v0 = __bfloat162float(f[(int64_t)d * hw]), v1 = __bfloat162float(f[(int64_t)(d + kCdSlices) * hw]);
      xa = fmaf(v0, v0, xa);
      xb = fmaf(v1, v1, xb);
    }
    if (d < D) {
      const float v0 = __bfloat162float(f[(int64_t)d * hw]);
      xa = fmaf(v0, v0, xa);
    }
    acc = xa + xb;
  }
  s_part[sl][px] = acc;
  __syncthreads();
  float xx = 0.f;
#pragma unroll
  for (int i = 0; i < kCdSlices; ++i) xx += s_part[i][px];  // same order in every slice: one value per pixel
  __syncthreads();
  float best = INFINITY;
  int arg = 0x7fffffff;
  if (live) {
    float* g = dist2 + (int64_t)b * Kc * hw + p;
    for (int k = sl; k < Kc; k += kCdSlices) {
      const float v = fmaxf(fmaf(-2.f, g[(int64_t)k * hw], xx + cc[k]), 0.f);
      g[(int64_t)k * hw] = v;
      if (v < best) {
        best = v;
        arg = k;
      }
    }
  }
  if (nearest == nullptr) return;
  s_part[sl][px] = best;
  s_arg[sl][px] = arg;
  __syncthreads();
  if (sl == 0 && live) {
#pragma unroll
    for (int i = 1; i < kCdSlices; ++i) {
      const float v = s_part[i][px];
      const int a = s_arg[i][px];
      if (v < best || (v == best && a < arg)) {
        best = v;
        arg = a;
      }
    }
    nearest[(int64_t)b * hw + p] = arg;
  }
}

}  // namespace bacs

using namespace bacs;

extern "C" {

size_t bacs_class_distance_workspace_bytes(int32_t B, int32_t Kc, int32_t D, int32_t h, int32_t w) {
#ifdef BACS_HAVE_CUTLASS
  if (B <= 0 || Kc <= 0 || D <= 0 || h <= 0 || w <= 0) return 0;
  const size_t cc = align_up((size_t)Kc * sizeof(float), 256);
  auto args = make_args(nullptr, nullptr, nullptr, h * w, Kc, D, B);
  return cc + align_up(Gemm::get_workspace_size(args), 256) + 256;
#else
  (void)B; (void)Kc; (void)D; (void)h; (void)w;
  return 0;
#endif
}

int bacs_class_distance(const void* features, int dtype, int32_t B, int32_t D, int32_t h, int32_t w, const void* protos,
                        int32_t Kc, float* dist2, int64_t* nearest, void* workspace, size_t workspace_bytes,
                        bacs_stream_t stream) {
#ifdef BACS_HAVE_CUTLASS
  BACS_REQUIRE(features && protos && dist2, "bacs_class_distance: null pointer");
  BACS_REQUIRE(dtype == BACS_BF16, "bacs_class_distance: bf16 features and prototypes only (tensor-core operands)");
  BACS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0 && Kc > 0 && Kc <= 256, "bacs_class_distance: bad shape (Kc <= 256)");
  const int hw = h * w;
  BACS_REQUIRE(hw % 8 == 0 && D % 8 == 0, "bacs_class_distance: h*w and D must be multiples of 8 (16-byte TMA rows)");
  BACS_REQUIRE((reinterpret_cast<uintptr_t>(features) & 15) == 0 && (reinterpret_cast<uintptr_t>(protos) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dist2) & 15) == 0,
               "bacs_class_distance: pointers must be 16-byte aligned");
  const size_t need = bacs_class_distance_workspace_bytes(B, Kc, D, h, w);
  if (!workspace || workspace_bytes < need) {
    set_error("bacs_class_distance: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return BACS_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  float* cc = reinterpret_cast<float*>(workspace);
  void* gemm_ws = reinterpret_cast<unsigned char*>(workspace) + align_up((size_t)Kc * sizeof(float), 256);
  class_sqnorm_kernel<<<(Kc + 7) / 8, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(protos), Kc, D, cc);
  BACS_CHECK_LAUNCH("bacs_class_distance(norms)");
  auto args = make_args(features, protos, dist2, hw, Kc, D, B);
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) {
    set_error("bacs_class_distance: the tensor-core GEMM cannot take this shape (hw=%d Kc=%d D=%d)", hw, Kc, D);
    return BACS_ERR_UNSUPPORTED;
  }
  if (gemm.initialize(args, gemm_ws, s) != cutlass::Status::kSuccess || gemm.run(s) != cutlass::Status::kSuccess) {
    set_error("bacs_class_distance: tensor-core GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return BACS_ERR_CUDA;
  }
  ++g_launch_count;
  dim3 grid((hw + kCdPx - 1) / kCdPx, B);
  class_distance_finish_kernel<<<grid, kCdPx * kCdSlices, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(features), D, hw, Kc, cc, dist2,
                                                    nearest);
  BACS_CHECK_LAUNCH("bacs_class_distance(finish)");
  return BACS_OK;
#else
  (void)features; (void)dtype; (void)B; (void)D; (void)h; (void)w; (void)protos; (void)Kc; (void)dist2; (void)nearest;
  (void)workspace; (void)workspace_bytes; (void)stream;
  set_error("bacs_class_distance: built without the CUTLASS header tree");
  return BACS_ERR_UNSUPPORTED;
#endif
}

}  // extern "C"
