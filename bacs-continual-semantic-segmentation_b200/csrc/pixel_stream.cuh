// Large class counts (K >= 64, ADE20K's 151): the per-pixel loss as two STREAMING passes over the logits.
//
// A [K][P] tile of K = 151 channels is 77 KB per 256 pixels: the shared-memory tile kernel (pixel_loss.cu) fits two
// single-stage CTAs per SM, only the CTA whose tile has landed computes, and it stays at 38 % of its HBM roofline
// whatever the tiling (profiles/r01_ncu_pixel_loss_ade_v25.txt).  Here nothing is staged:
//   pass A  reads every logit once (16-byte vectors of adjacent pixels, channels strided by H*W, 4 loads in flight per
//           thread), keeps an ONLINE soft-max per pixel (running max, old / new class exponent sums rescaled when a
//           chunk raises the max, first arg-max), evaluates the same pixel_terms() as the other kernels and writes six
//           fp32 coefficients per pixel (24 B) + arg-max + distill mask;
//   pass B  reads the logits a second time and writes the gradient  e_k * cg[group(k)] - [k==0] d0 - [k==y] dy.
// HBM traffic is 3 K s + 64 bytes per pixel instead of the algorithmic 2 K s + 17 (one extra read of the logits), but
// both passes are plain coalesced streams.  Measured at B=24, 512x512, K=151 (one B200): pass B 717 us (5.5 TB/s),
// pass A 610 us with four pixels per thread (80 registers, three CTAs per SM; 764 us with eight pixels and two CTAs,
// 43 % of the warp samples waiting on a just-issued load): issue slots 60 % busy, DRAM 44 %;
// fp32 logits 2.08 ms against 2.90 ms for the tile kernel, bf16 1.35 ms against 1.57 ms.
#pragma once
#include "pixel_common.cuh"

namespace bacs {

struct alignas(16) StreamParams {
  bacs_pixel_args a;
  float* coef;        // [6][B*H*W]: nm = -max*log2e, cg0, cg1, cg2, d0, dy   (nullptr: no gradient pass)
  double* partials;   // [B * blocks_x][BACS_NACC]
  int blocks_x;       // blocks per image
  int b0;             // first image of this launch
  float inv_n, sy, sx;
};

template <typename T> struct StreamVec;
template <> struct StreamVec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct StreamVec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct StreamVec<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <int BYTES> struct StreamWord;
template <> struct StreamWord<16> { using type = uint4; };
template <> struct StreamWord<8> { using type = uint2; };
template <> struct StreamWord<4> { using type = uint32_t; };

constexpr int kStreamThreads = 256;
constexpr int kStreamStatsPx = 4;  // pixels per thread of the statistics pass (8- or 16-byte loads): its state fits 80 registers
constexpr int kStreamChunk = 4;  // channels (16-byte loads) in flight per thread

// ---- per-pixel stage (everything that follows the soft-max statistics of N adjacent pixels p0 .. p0+N-1 of image b):
// labels, seen probability from the T head maps, pixel_terms(), distill mask, focal gradient, arg-max.  The gradient
// coefficients of pixel j leave through emit(j, -max*log2e, coefficients, label).
// PRE (N == 1 only): the caller has already loaded the label (pre_lab) and the label's own logit (pre_xy).
template <typename T, int N, bool PRE = false, typename Emit>
__device__ __forceinline__ void stream_pixel_stage(const StreamParams& p, int b, float s_norm, int old_cl, bool have_seen,
                                                   int64_t p0, const T* base, const float (&m)[N], const float (&so)[N],
                                                   const float (&sn)[N], const float (&x0)[N], const int (&am)[N],
                                                   float* acc, Emit emit, long long pre_lab = 0, float pre_xy = 0.f) {
  static_assert(!PRE || N == 1, "preloaded labels: one pixel per thread");
  const bacs_pixel_args& a = p.a;
  const int K = a.K;
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t pix0 = (int64_t)b * HW + p0;
  long long lab[N];
  {
    const int64_t* lp = a.labels + pix0;
    if constexpr (PRE) {
      lab[0] = pre_lab;
    } else if constexpr (N == 1) {
      lab[0] = __ldg(lp);
    } else {
#pragma unroll
      for (int j = 0; j + 1 < N; j += 2) {
        const longlong2 t = __ldg(reinterpret_cast<const longlong2*>(lp + j));
        lab[j] = t.x;
        lab[j + 1] = t.y;
      }
    }
  }
  uint32_t mbits = 0;
  // focal gradient: runs of pixels that share the low-res cell are merged before the atomics
  int run_cell = -1, run_dx = 0, run_dy = 0;
  float r00 = 0.f, r01 = 0.f, r10 = 0.f, r11 = 0.f;
  float* gzb = a.gz ? a.gz + (int64_t)b * a.h * a.w : nullptr;
  auto run_flush = [&]() {
    if (run_cell >= 0) {
      if (r00 != 0.f) atomicAdd(gzb + run_cell, r00);
      if (r01 != 0.f) atomicAdd(gzb + run_cell + run_dx, r01);
      if (r10 != 0.f) atomicAdd(gzb + run_cell + run_dy, r10);
      if (r11 != 0.f) atomicAdd(gzb + run_cell + run_dy + run_dx, r11);
    }
    run_cell = -1;
    r00 = r01 = r10 = r11 = 0.f;
  };
#pragma unroll
  for (int j = 0; j < N; ++j) {
    int y = -1;
    bool is_ign = true;
    const long long l = lab[j];
    if (l == a.ignore_index) {
    } else if (l >= 0 && l < K) {
      y = (int)l;
      is_ign = false;
    } else {
      acc[BACS_ACC_INVALID] += 1.f;
    }
    const float nmj = -m[j] * kLog2e;
    const float e0 = ex2_fast(fmaf(x0[j], kLog2e, nmj));
    const float S_fg = so[j] + sn[j];
    const float S = S_fg + e0;
    const float S_old = so[j] + (old_cl >= 1 ? e0 : 0.f);
    float xy;
    if constexpr (PRE) xy = y > 0 ? pre_xy : x0[j];
    else xy = y > 0 ? DT<T>::to_f(base[(int64_t)y * HW + j]) : x0[j];
    float seen = 0.f, zfoc = 0.f, wx1 = 0.f, wy1 = 0.f;
    int cell = -1, cdx = 0, cdy = 0;
    if (a.seen_max) seen = __ldg(a.seen_max + pix0 + j);
    if (a.z) {
      const int64_t pix = p0 + j;
      const int Y = (int)(pix / a.W), X = (int)(pix - (int64_t)Y * a.W);
      const Lerp ly = lerp_align_corners(Y, a.h, p.sy), lx = lerp_align_corners(X, a.w, p.sx);
      const float wx0 = 1.f - lx.w1, wy0 = 1.f - ly.w1;
      const float* zb = a.z + (int64_t)b * a.T * a.h * a.w;
      const int o00 = ly.i0 * a.w + lx.i0, o01 = ly.i0 * a.w + lx.i1;
      const int o10 = ly.i1 * a.w + lx.i0, o11 = ly.i1 * a.w + lx.i1;
      float zmax = -INFINITY;
      for (int t = 0; t < a.T; ++t) {
        const float* zt = zb + (int64_t)t * a.h * a.w;
        const float left = __fadd_rn(__fmul_rn(wy0, __ldg(zt + o00)), __fmul_rn(ly.w1, __ldg(zt + o10)));
        const float right = __fadd_rn(__fmul_rn(wy0, __ldg(zt + o01)), __fmul_rn(ly.w1, __ldg(zt + o11)));
        const float v = __fadd_rn(__fmul_rn(wx0, left), __fmul_rn(lx.w1, right));
        zmax = fmaxf(zmax, v);
        if (t == a.focal_head) zfoc = v;
      }
      cell = o00;
      cdx = lx.i1 - lx.i0;
      cdy = (ly.i1 - ly.i0) * a.w;
      wx1 = lx.w1;
      wy1 = ly.w1;
      if (!a.seen_max) seen = sigmoid_fast(zmax);
    }
    PixCoef pc;
    float gfoc;
    uint8_t dm;
    pixel_terms(a, p.inv_n, s_norm, old_cl, y, is_ign, m[j], S, S_old, S_fg, e0, x0[j], xy, seen, have_seen, zfoc, acc,
                pc, gfoc, dm);
    mbits |= (uint32_t)dm << (8 * (j & 3));
    if (N == 1) {
      if (a.distill_mask) a.distill_mask[pix0] = dm;
    } else if (N == 2) {
      if (j == 1 && a.distill_mask) *reinterpret_cast<uint16_t*>(a.distill_mask + pix0) = (uint16_t)mbits;
    } else if ((j & 3) == 3) {
      if (a.distill_mask) *reinterpret_cast<uint32_t*>(a.distill_mask + pix0 + (j & ~3)) = mbits;
      mbits = 0;
    }
    emit(j, nmj, pc, l);
    if (a.gz) {
      if (cell != run_cell || cdx != run_dx || cdy != run_dy) run_flush();
      if (gfoc != 0.f) {
        run_cell = cell;
        run_dx = cdx;
        run_dy = cdy;
        const float t0 = gfoc * (1.f - wy1), t1 = gfoc * wy1;
        r00 = fmaf(t0, 1.f - wx1, r00);
        r01 = fmaf(t0, wx1, r01);
        r10 = fmaf(t1, 1.f - wx1, r10);
        r11 = fmaf(t1, wx1, r11);
      }
    }
  }
  if (a.gz) run_flush();
  if (a.preds) {
    int64_t* out = a.preds + pix0;
    if constexpr (N == 1) {
      out[0] = (long long)am[0];
    } else {
#pragma unroll
      for (int j = 0; j + 1 < N; j += 2)
        *reinterpret_cast<longlong2*>(out + j) = make_longlong2((long long)am[j], (long long)am[j + 1]);
    }
  }
}

// ---- pass A: statistics, loss terms, coefficients, arg-max, distill mask, focal gradient ---------------------------
// groups of kStreamStatsPx pixels g_first, g_first + g_step, ... < g_end of image b
template <typename T>
__device__ __forceinline__ void stream_stats_groups(const StreamParams& p, int b, float s_norm, int64_t g_first, int64_t g_end,
                                                    int64_t g_step, float* acc) {
  constexpr int N = kStreamStatsPx, CH = kStreamChunk;
  using vec_t = typename StreamWord<sizeof(T) * N>::type;
  const bacs_pixel_args& a = p.a;
  const int K = a.K;
  const int64_t HW = (int64_t)a.H * a.W, NPIX = HW * a.B;
  const int old_cl = min(max(a.old_cl, 0), K);
  const bool have_seen = (a.z != nullptr) || (a.seen_max != nullptr);
  const T* img = reinterpret_cast<const T*>(a.logits) + (int64_t)b * K * HW;
  for (int64_t g = g_first; g < g_end; g += g_step) {
    const int64_t p0 = g * N;
    const T* base = img + p0;
    // Running max / first arg-max on the PACKED storage words (HMNMX2 + HSET2 + LOP3 per pixel pair, as in the tile
    // kernels); the exponent sums are taken against a per-pixel REFERENCE r <= running max that is only moved up when
    // the max has run more than 32 ahead of it (checked every 8 channels), so the common path has no rescaling.
    constexpr int W = N / 2;
    using R = Raw<T>;
    using reg_t = typename R::reg_t;
    static_assert(sizeof(reg_t) * W == sizeof(vec_t), "a vector is W packed pixel pairs");
    float m[N], so[N], sn[N], x0[N];
    F2 so2[W], sn2[W], nr2[W];  // packed pixel pairs: exponent sums, nr = -reference * log2e
    int am[N];
    typename R::Max mt[W];
    {
      const vec_t r4 = __ldg(reinterpret_cast<const vec_t*>(base));
      const reg_t* rw = reinterpret_cast<const reg_t*>(&r4);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        mt[w] = R::init(rw[w]);
        R::set_first(mt[w], 0);
        R::unpack(rw[w], x0[2 * w], x0[2 * w + 1]);
      }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      nr2[w] = f2(-x0[2 * w] * kLog2e, -x0[2 * w + 1] * kLog2e);
      so2[w] = sn2[w] = f2b(0.f);
    }
    const F2 l2e2 = f2b(kLog2e);
    // a block of channels is first folded into the running max, then the reference is lifted if the max has run ahead,
    // and only then are the exponents taken: no logit of the block is more than 32 above the reference (no overflow,
    // whatever the jumps between neighbouring channels)
    auto track = [&](const vec_t& r4, int c) {
      const reg_t* rw = reinterpret_cast<const reg_t*>(&r4);
#pragma unroll
      for (int w = 0; w < W; ++w) R::update(mt[w], rw[w], c);
    };
    auto accumulate = [&](const vec_t& r4, F2* sum) {
      const reg_t* rw = reinterpret_cast<const reg_t*>(&r4);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        float v0, v1;
        R::unpack(rw[w], v0, v1);
        const F2 arg = fma2(f2(v0, v1), l2e2, nr2[w]);
        sum[w] = add2(sum[w], f2(ex2_fast(f2lo(arg)), ex2_fast(f2hi(arg))));
      }
    };
    auto lift_reference = [&]() {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        float m0, m1;
        int a0, a1;
        R::finish(mt[w], m0, m1, a0, a1);
        const float a0h = fmaf(m0, kLog2e, f2lo(nr2[w])), a1h = fmaf(m1, kLog2e, f2hi(nr2[w]));  // (max - ref) * log2e
        const bool l0 = a0h > 32.f * kLog2e, l1 = a1h > 32.f * kLog2e;
        if (l0 || l1) {
          const F2 f = f2(l0 ? ex2_fast(-a0h) : 1.f, l1 ? ex2_fast(-a1h) : 1.f);
          so2[w] = mul2(so2[w], f);
          sn2[w] = mul2(sn2[w], f);
          nr2[w] = f2(l0 ? -m0 * kLog2e : f2lo(nr2[w]), l1 ? -m1 * kLog2e : f2hi(nr2[w]));
        }
      }
    };
    auto range = [&](int cbeg, int cend, F2* sum) {
      int c = cbeg;
      for (; c + 2 * CH <= cend; c += 2 * CH) {  // 8 independent 16-byte loads in flight per thread
        vec_t raw[2 * CH];
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) raw[i] = __ldg(reinterpret_cast<const vec_t*>(base + (int64_t)(c + i) * HW));
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) track(raw[i], c + i);
        lift_reference();
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) accumulate(raw[i], sum);
      }
      for (; c < cend; ++c) {
        const vec_t r4 = __ldg(reinterpret_cast<const vec_t*>(base + (int64_t)c * HW));
        track(r4, c);
        lift_reference();
        accumulate(r4, sum);
      }
    };
    range(1, old_cl, so2);
    range(max(old_cl, 1), K, sn2);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      int a0, a1;
      R::finish(mt[w], m[2 * w], m[2 * w + 1], a0, a1);
      am[2 * w] = a0;
      am[2 * w + 1] = a1;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {  // sums against the reference -> sums against the max
      const float nrj = (j & 1) ? f2hi(nr2[j >> 1]) : f2lo(nr2[j >> 1]);
      const float f = ex2_fast(-fmaf(m[j], kLog2e, nrj));
      so[j] = f * ((j & 1) ? f2hi(so2[j >> 1]) : f2lo(so2[j >> 1]));
      sn[j] = f * ((j & 1) ? f2hi(sn2[j >> 1]) : f2lo(sn2[j >> 1]));
    }
    float* cf = p.coef ? p.coef + (int64_t)b * HW + p0 : nullptr;
    stream_pixel_stage<T, N>(p, b, s_norm, old_cl, have_seen, p0, base, m, so, sn, x0, am, acc,
                             [cf, NPIX](int j, float nmj, const PixCoef& pc, long long) {
                               if (cf) {  // (a thread's N pixels fill whole 32-byte sectors of every coefficient plane)
                                 cf[0 * NPIX + j] = nmj;
                                 cf[1 * NPIX + j] = pc.cg0;
                                 cf[2 * NPIX + j] = pc.cg1;
                                 cf[3 * NPIX + j] = pc.cg2;
                                 cf[4 * NPIX + j] = pc.d0;
                                 cf[5 * NPIX + j] = pc.dy;
                               }
                             });
  }
}

// CE-type gradient normaliser 1 / sum_c hist[c] w[c] (0 for WEIGHTED_CE); every thread of the block gets it
__device__ __forceinline__ float stream_norm(const StreamParams& p, float* s_norm_sh) {
  const bacs_pixel_args& a = p.a;
  const int tid = threadIdx.x;
  if (tid == 0) *s_norm_sh = 0.f;
  __syncthreads();
  if (tid < 32 && a.mode != BACS_PIX_WEIGHTED_CE && p.coef != nullptr) {
    double s = 0.0;
    for (int c = tid; c < a.K && c < 256; c += 32)
      if (c != a.ignore_index)
        s += (double)a.hist[c] * ((a.mode == BACS_PIX_CE && a.class_w) ? (double)a.class_w[c] : 1.0);
    s = warp_sum(s);
    if (tid == 0) *s_norm_sh = s > 0.0 ? (float)(1.0 / s) : 0.f;
  }
  __syncthreads();
  return *s_norm_sh;
}

// per-CTA partial sums -> p.partials[slot]
__device__ __forceinline__ void stream_flush_acc(const StreamParams& p, const float* acc, float (*red_scratch)[BACS_NACC],
                                                 int64_t slot, int nwarps = kStreamThreads / 32) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int i = 0; i < BACS_NACC; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red_scratch[wid][i] = v;
  }
  __syncthreads();
  if (tid < BACS_NACC) {
    double v = 0.0;
    for (int wv = 0; wv < nwarps; ++wv) v += (double)red_scratch[wv][tid];
    p.partials[slot * BACS_NACC + tid] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(kStreamThreads, 3) pixel_stream_stats_kernel(const __grid_constant__ StreamParams p) {
  __shared__ float red_scratch[kStreamThreads / 32][BACS_NACC];
  __shared__ float s_norm_sh;
  const int b = p.b0 + (int)blockIdx.y;
  const float s_norm = stream_norm(p, &s_norm_sh);
  float acc[BACS_NACC];
#pragma unroll
  for (int i = 0; i < BACS_NACC; ++i) acc[i] = 0.f;
  const int64_t groups = (int64_t)p.a.H * p.a.W / kStreamStatsPx;
  stream_stats_groups<T>(p, b, s_norm, (int64_t)blockIdx.x * kStreamThreads + threadIdx.x, groups,
                         (int64_t)gridDim.x * kStreamThreads, acc);
  stream_flush_acc(p, acc, red_scratch, (int64_t)b * gridDim.x + blockIdx.x);
}

// ---- pass B: gradient ----------------------------------------------------------------------------------------------
// N pixels of a thread as 32-bit words of the storage type
template <typename T, int N> struct StreamIO {
  static constexpr int WORDS = (int)(sizeof(T) * N / 4);
  using vec_t = typename StreamWord<sizeof(T) * N>::type;
  __device__ static __forceinline__ void unpack(const vec_t& r, float* v) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = __uint_as_float(w[i]);
    } else {
#pragma unroll
      for (int i = 0; i < WORDS; ++i) Raw<T>::unpack(w[i], v[2 * i], v[2 * i + 1]);
    }
  }
  __device__ static __forceinline__ vec_t pack(const float* v) {
    vec_t r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < N; ++i) w[i] = __float_as_uint(v[i]);
    } else if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int i = 0; i < WORDS; ++i) {
        if constexpr (DT<T>::id == BACS_BF16) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          w[i] = *reinterpret_cast<const uint32_t*>(&h);
        } else {
          const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
          w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
    }
    return r;
  }
};

// groups of N pixels g_first, g_first + g_step, ... < g_end of image b.
// (measured and not used: statistics and gradient of a 1024-pixel strip back to back in one persistent
//  launch, channels walked backwards with streaming hints so that the second read is served by L2: DRAM reads fall from
//  3.85 to 2.73 GB per launch at K = 151 but the launch takes 1.37 ms against 1.34 ms for the two whole-batch passes; both
//  are bound by issue / latency at 16 warps per SM, not by HBM.)
template <typename T, int N>
__device__ __forceinline__ void stream_grad_groups(const StreamParams& p, int b, int64_t g_first, int64_t g_end, int64_t g_step) {
  constexpr int CH = kStreamChunk;
  using IO = StreamIO<T, N>;
  using vec_t = typename IO::vec_t;
  const bacs_pixel_args& a = p.a;
  const int K = a.K;
  const int64_t HW = (int64_t)a.H * a.W, NPIX = HW * a.B;
  const int old_cl = min(max(a.old_cl, 0), K);
  const T* img = reinterpret_cast<const T*>(a.logits) + (int64_t)b * K * HW;
  T* gimg = reinterpret_cast<T*>(a.dlogits) + (int64_t)b * K * HW;
  for (int64_t g = g_first; g < g_end; g += g_step) {
    const int64_t p0 = g * N, pix0 = (int64_t)b * HW + p0;
    const float* cf = p.coef + pix0;
    float nm[N], cg0[N], cg1[N], cg2[N], d0[N], dy[N];
#pragma unroll
    for (int j = 0; j < N; j += 4) {
      const float4 t0 = *reinterpret_cast<const float4*>(cf + 0 * NPIX + j), t1 = *reinterpret_cast<const float4*>(cf + 1 * NPIX + j);
      const float4 t2 = *reinterpret_cast<const float4*>(cf + 2 * NPIX + j), t3 = *reinterpret_cast<const float4*>(cf + 3 * NPIX + j);
      const float4 t4 = *reinterpret_cast<const float4*>(cf + 4 * NPIX + j), t5 = *reinterpret_cast<const float4*>(cf + 5 * NPIX + j);
      nm[j] = t0.x; nm[j + 1] = t0.y; nm[j + 2] = t0.z; nm[j + 3] = t0.w;
      cg0[j] = t1.x; cg0[j + 1] = t1.y; cg0[j + 2] = t1.z; cg0[j + 3] = t1.w;
      cg1[j] = t2.x; cg1[j + 1] = t2.y; cg1[j + 2] = t2.z; cg1[j + 3] = t2.w;
      cg2[j] = t3.x; cg2[j + 1] = t3.y; cg2[j + 2] = t3.z; cg2[j + 3] = t3.w;
      d0[j] = t4.x; d0[j + 1] = t4.y; d0[j + 2] = t4.z; d0[j + 3] = t4.w;
      dy[j] = t5.x; dy[j + 1] = t5.y; dy[j + 2] = t5.z; dy[j + 3] = t5.w;
    }
    const T* base = img + p0;
    T* gbase = gimg + p0;
    const int nchunk = (K + CH - 1) / CH;
    for (int q = 0; q < nchunk; ++q) {
      const int c0 = q * CH;
      vec_t raw[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (c0 + i < K) {
          const vec_t* src = reinterpret_cast<const vec_t*>(base + (int64_t)(c0 + i) * HW);
          raw[i] = __ldg(src);
        }
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c = c0 + i;
        if (c < K) {
          float v[N], gq[N];
          IO::unpack(raw[i], v);
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const float e = ex2_fast(fmaf(v[j], kLog2e, nm[j]));
            gq[j] = c == 0 ? fmaf(e, cg0[j], -d0[j]) : e * (c < old_cl ? cg1[j] : cg2[j]);
          }
          vec_t* dst = reinterpret_cast<vec_t*>(gbase + (int64_t)c * HW);
          *dst = IO::pack(gq);
        }
      }
    }
    // the label's own channel: recomputed in fp32 so that -dy is applied before the rounding to the storage type
    bool any = false;
#pragma unroll
    for (int j = 0; j < N; ++j) any |= dy[j] != 0.f;
    if (any) {
      const int64_t* lp = a.labels + pix0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        if (dy[j] != 0.f) {
          const int y = (int)lp[j];
          const float x = DT<T>::to_f(base[(int64_t)y * HW + j]);
          const float e = ex2_fast(fmaf(x, kLog2e, nm[j]));
          const float cg = y == 0 ? cg0[j] : (y < old_cl ? cg1[j] : cg2[j]);
          gbase[(int64_t)y * HW + j] = DT<T>::from_f(e * cg - dy[j] - (y == 0 ? d0[j] : 0.f));
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kStreamThreads, 2) pixel_stream_grad_kernel(const __grid_constant__ StreamParams p) {
  constexpr int N = StreamVec<T>::N;
  const int b = p.b0 + (int)blockIdx.y;
  const int64_t groups = (int64_t)p.a.H * p.a.W / N;
  stream_grad_groups<T, N>(p, b, (int64_t)blockIdx.x * kStreamThreads + threadIdx.x, groups,
                                  (int64_t)gridDim.x * kStreamThreads);
}

}  // namespace bacs
