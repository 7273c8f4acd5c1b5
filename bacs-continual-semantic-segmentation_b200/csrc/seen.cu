// Seen / unseen detector heads at feature resolution.
//   reference: networks/bg_detector.py:17-40 (classification_head.get_distance / predict),
//              100-165 (get_seen_map_task, forward_seen_before, get_seen_probs)
// z[b,t,q] = bias_t + sum_c w[t,c] * |sigmoid(f[b,c,q]) - sigmoid(proto[t,c])|
// A weighted L1 distance per task head: HBM/ALU work, no GEMM form.
#include "common.cuh"

namespace bacs {

constexpr int kSeenWarps = 8;
constexpr int kSeenPix = 2;  // pixels per thread: the per-channel weights are fetched once for all of them

// block = 32 lanes x 8 channel groups; a block covers 128 consecutive pixels of one image
template <typename T, int TMAX>
__global__ void __launch_bounds__(32 * kSeenWarps) seen_logits_kernel(const T* __restrict__ feat, int D, int hw,
                                                                       const float* __restrict__ proto,
                                                                       const float* __restrict__ weight,
                                                                       const float* __restrict__ bias, int Tn,
                                                                       int chunk, float* __restrict__ z) {
  extern __shared__ float smem[];
  float* s_sp = smem;              // [chunk][Tn] sigmoid(proto)   (head index fastest: one row per channel)
  float* s_w = smem + Tn * chunk;  // [chunk][Tn]
  const int lane = threadIdx.x, cg = threadIdx.y;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * (32 * kSeenPix) + lane;
  const int tid = cg * 32 + lane;
  float acc[kSeenPix][TMAX];
#pragma unroll
  for (int i = 0; i < kSeenPix; ++i)
#pragma unroll
    for (int t = 0; t < TMAX; ++t) acc[i][t] = 0.f;
  const T* base = feat + (int64_t)b * D * hw;
  for (int c0 = 0; c0 < D; c0 += chunk) {
    const int cn = min(chunk, D - c0);
    __syncthreads();
    for (int i = tid; i < Tn * cn; i += 32 * kSeenWarps) {
      const int t = i / cn, c = i - t * cn;
      s_sp[c * Tn + t] = sigmoid_fast(proto[t * D + c0 + c]);
      s_w[c * Tn + t] = weight[t * D + c0 + c];
    }
    __syncthreads();
    // warp cg handles channels cg, cg + 8, ... of the chunk; two channels (8 loads) in flight per thread
    for (int c = cg; c < cn; c += 2 * kSeenWarps) {
      float x[2][kSeenPix];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int cc = c + u * kSeenWarps;
#pragma unroll
        for (int i = 0; i < kSeenPix; ++i) {
          const int q = q0 + 32 * i;
          x[u][i] = (cc < cn && q < hw) ? DT<T>::to_f(base[(int64_t)(c0 + cc) * hw + q]) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int cc = c + u * kSeenWarps;
        if (cc < cn) {
          float sx[kSeenPix];
#pragma unroll
          for (int i = 0; i < kSeenPix; ++i) sx[i] = sigmoid_fast(x[u][i]);
#pragma unroll
          for (int t = 0; t < TMAX; ++t)
            if (t < Tn) {
              const float wv = s_w[cc * Tn + t], sp = s_sp[cc * Tn + t];
#pragma unroll
              for (int i = 0; i < kSeenPix; ++i) acc[i][t] = fmaf(wv, fabsf(sx[i] - sp), acc[i][t]);
            }
        }
      }
    }
  }
  __syncthreads();
  // cross-warp reduction through shared memory: red[cg][t][pixel slot]
  float* red = smem;
  constexpr int PS = 32 * kSeenPix;
#pragma unroll
  for (int t = 0; t < TMAX; ++t)
    if (t < Tn) {
#pragma unroll
      for (int i = 0; i < kSeenPix; ++i) red[(cg * Tn + t) * PS + 32 * i + lane] = acc[i][t];
    }
  __syncthreads();
  for (int t = cg; t < Tn; t += kSeenWarps) {
#pragma unroll
    for (int i = 0; i < kSeenPix; ++i) {
      float s = bias[t];
#pragma unroll
      for (int g = 0; g < kSeenWarps; ++g) s += red[(g * Tn + t) * PS + 32 * i + lane];
      const int q = q0 + 32 * i;
      if (q < hw) z[((int64_t)b * Tn + t) * hw + q] = s;
    }
  }
}

__global__ void __launch_bounds__(256) seen_upsample_kernel(const float* __restrict__ z, int BT, int h, int w, int H,
                                                            int W, float sy, float sx, int apply_sigmoid,
                                                            float* __restrict__ out) {
  const int64_t total = (int64_t)BT * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const int64_t bt = i / ((int64_t)W * H);
    const Lerp ly = lerp_align_corners(Y, h, sy), lx = lerp_align_corners(X, w, sx);
    const float* p = z + bt * h * w;
    const float v00 = p[ly.i0 * w + lx.i0], v01 = p[ly.i0 * w + lx.i1];
    const float v10 = p[ly.i1 * w + lx.i0], v11 = p[ly.i1 * w + lx.i1];
    const float wx0 = 1.f - lx.w1, wy0 = 1.f - ly.w1;
    // same operation order as ATen's upsample_bilinear2d
    const float top = __fadd_rn(__fmul_rn(wx0, v00), __fmul_rn(lx.w1, v01));
    const float bot = __fadd_rn(__fmul_rn(wx0, v10), __fmul_rn(lx.w1, v11));
    float v = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(ly.w1, bot));
    if (apply_sigmoid) v = sigmoid_acc(v);
    out[i] = v;
  }
}

// One block per channel: dW[c] = s * sum_{b,q} gz * |sig(f) - sig(p)|, optional dfeat.
template <typename T>
__global__ void __launch_bounds__(512) seen_head_backward_kernel(const T* __restrict__ feat, int B, int D, int hw,
                                                                 const float* __restrict__ proto_t,
                                                                 const float* __restrict__ weight_t,
                                                                 const float* __restrict__ gz,
                                                                 const float* __restrict__ scale_dev,
                                                                 float* __restrict__ dweight, float* __restrict__ dbias,
                                                                 T* __restrict__ dfeat) {
  __shared__ float scratch[32];
  const int c = blockIdx.x;
  const float scale = scale_dev ? *scale_dev : 1.f;
  if (c == D) {  // extra block: bias gradient
    float s = 0.f;
    for (int i = threadIdx.x; i < B * hw; i += blockDim.x) s += gz[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) dbias[0] = s * scale;
    return;
  }
  const float sp = sigmoid_fast(proto_t[c]);
  const float wc = weight_t[c];
  float acc = 0.f;
  // threads are split over images (ib) and pixels (iq) so that no index needs a division
  constexpr int U = 4;
  const int tq = min((int)blockDim.x, ((hw + U - 1) / U + 31) / 32 * 32);  // threads along the pixel axis
  const int nb_par = max(1, (int)blockDim.x / tq);                          // images processed side by side
  const int ib = threadIdx.x / tq, iq = threadIdx.x - ib * tq;
  if (ib < nb_par) {
    for (int b = ib; b < B; b += nb_par) {
      const T* row = feat + ((int64_t)b * D + c) * hw;
      const float* g = gz + (int64_t)b * hw;
      T* drow = dfeat ? dfeat + ((int64_t)b * D + c) * hw : nullptr;
      for (int q0 = iq; q0 < hw; q0 += tq * U) {
        float x[U], gq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int q = q0 + u * tq;
          x[u] = q < hw ? DT<T>::to_f(row[q]) : 0.f;
          gq[u] = q < hw ? g[q] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int q = q0 + u * tq;
          if (q < hw) {
            const float sx = sigmoid_fast(x[u]);
            const float d = sx - sp;
            acc = fmaf(gq[u], fabsf(d), acc);
            if (drow) {
              const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
              drow[q] = DT<T>::from_f(scale * gq[u] * wc * sg * sx * (1.f - sx));
            }
          }
        }
      }
    }
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) dweight[c] = acc * scale;
}

// scale = weight * ready * [#bg > 0] / #kept ; focal = scale * sum(focal terms)
__global__ void focal_scale_kernel(const double* __restrict__ acc, const int32_t* __restrict__ ready, float weight,
                                   float* __restrict__ scale_out, double* __restrict__ out2) {
  const double kept = acc[BACS_ACC_KEPT];
  const bool on = (ready == nullptr || *ready != 0) && acc[BACS_ACC_BG] > 0.0 && kept > 0.0;
  const double s = on ? (double)weight / kept : 0.0;
  if (scale_out) *scale_out = (float)s;
  if (out2) {
    out2[0] = s;
    out2[1] = s * acc[BACS_ACC_FOCAL];
  }
}

}  // namespace bacs

using namespace bacs;

extern "C" {

int bacs_seen_logits(const void* features, int dtype, int B, int D, int h, int w, const float* proto,
                     const float* weight, const float* bias, int T, float* z, bacs_stream_t stream) {
  BACS_REQUIRE(features && proto && weight && bias && z, "bacs_seen_logits: null pointer");
  BACS_REQUIRE(B > 0 && B < 65536 && D > 0 && h > 0 && w > 0, "bacs_seen_logits: bad shape");
  BACS_REQUIRE(T > 0 && T <= 32, "bacs_seen_logits: T=%d not in [1,32]", T);
  const int hw = h * w;
  // chunk of channels whose sigmoid(proto) / weight rows fit in <= 48 KB of shared memory
  int chunk = (48 * 1024 / 4 / 2) / T;
  chunk = chunk / kSeenWarps * kSeenWarps;
  if (chunk > D) chunk = (D + kSeenWarps - 1) / kSeenWarps * kSeenWarps;
  size_t smem = (size_t)2 * T * chunk * sizeof(float);
  const size_t red = (size_t)kSeenWarps * T * 32 * kSeenPix * sizeof(float);
  if (smem < red) smem = red;
  dim3 grid((hw + 32 * kSeenPix - 1) / (32 * kSeenPix), B), block(32, kSeenWarps);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_Z(TT, TM)                                                                                          \
  do {                                                                                                            \
    auto kern = seen_logits_kernel<TT, TM>;                                                                       \
    if (smem > 48 * 1024) {                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
      if (e != cudaSuccess) {                                                                                     \
        set_error("bacs_seen_logits: shared memory opt-in failed: %s", cudaGetErrorString(e));                   \
        return BACS_ERR_CUDA;                                                                                     \
      }                                                                                                           \
    }                                                                                                             \
    kern<<<grid, block, smem, s>>>(reinterpret_cast<const TT*>(features), D, hw, proto, weight, bias, T, chunk, z); \
  } while (0)
  BACS_DISPATCH_DTYPE(dtype, TT, {
    if (T <= 4) LAUNCH_Z(TT, 4);
    else if (T <= 8) LAUNCH_Z(TT, 8);
    else if (T <= 16) LAUNCH_Z(TT, 16);
    else LAUNCH_Z(TT, 32);
  });
#undef LAUNCH_Z
  BACS_CHECK_LAUNCH("bacs_seen_logits");
  return BACS_OK;
}

int bacs_seen_upsample(const float* z, int B, int T, int h, int w, int scale, int apply_sigmoid, float* out,
                       bacs_stream_t stream) {
  BACS_REQUIRE(z && out, "bacs_seen_upsample: null pointer");
  BACS_REQUIRE(B > 0 && T > 0 && h > 0 && w > 0 && scale > 0, "bacs_seen_upsample: bad shape");
  const int H = h * scale, W = w * scale;
  const int64_t total = (int64_t)B * T * H * W;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  seen_upsample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, B * T, h, w, H, W, ac_scale(h, H),
                                                                           ac_scale(w, W), apply_sigmoid, out);
  BACS_CHECK_LAUNCH("bacs_seen_upsample");
  return BACS_OK;
}

int bacs_seen_head_backward(const void* features, int dtype, int B, int D, int h, int w, const float* proto_t,
                            const float* weight_t, const float* gz, const float* scale_dev, float* dweight,
                            float* dbias, void* dfeatures, bacs_stream_t stream) {
  BACS_REQUIRE(features && proto_t && weight_t && gz && dweight && dbias, "bacs_seen_head_backward: null pointer");
  BACS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, "bacs_seen_head_backward: bad shape");
  const int hw = h * w;
  cudaStream_t s = (cudaStream_t)stream;
  BACS_DISPATCH_DTYPE(dtype, TT, {
    seen_head_backward_kernel<TT><<<D + 1, 512, 0, s>>>(reinterpret_cast<const TT*>(features), B, D, hw, proto_t,
                                                        weight_t, gz, scale_dev, dweight, dbias,
                                                        reinterpret_cast<TT*>(dfeatures));
  });
  BACS_CHECK_LAUNCH("bacs_seen_head_backward");
  return BACS_OK;
}

int bacs_focal_scale(const double* acc, const int32_t* ready, float weight, float* scale_out, double* out2,
                     bacs_stream_t stream) {
  BACS_REQUIRE(acc, "bacs_focal_scale: null pointer");
  focal_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(acc, ready, weight, scale_out, out2);
  BACS_CHECK_LAUNCH("bacs_focal_scale");
  return BACS_OK;
}

}  // extern "C"
