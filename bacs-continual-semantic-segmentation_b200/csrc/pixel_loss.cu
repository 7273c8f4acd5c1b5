// The fused per-pixel kernel of the BACS loss step.
//
// One CTA owns a tile of P consecutive pixels of one image.  The K logit rows of the tile
// ([K][P], NCHW so each row is contiguous) are staged in shared memory by 1-D TMA bulk
// copies (cp.async.bulk + mbarrier); labels, the low-res seen logits and the bilinear taps
// are fetched into registers while the copies are in flight.  Softmax statistics are then
// computed once per pixel and feed, in the same pass:
//   * background-aware unbiased CE  (training/loss_utils.py:542-585)   mode WEIGHTED_CE
//   * plain / class-weighted CE     (loss/base_loss.py:237-240)        mode CE
//   * MiB unbiased CE               (training/loss_utils.py:492-520)   mode UNBIASED_CE
//   * per-image importance score    (loss/bacs_loss.py:183-189)        mode SCORE
//   * seen-detector focal loss of one head + its gradient w.r.t. the low-res head output
//     (loss/base_loss.py:255-272, smp FocalLoss binary; bilinear x16 align_corners=True
//      evaluated on the fly, networks/bg_detector.py:13-15)
//   * the teacher-distill pixel mask (loss/bacs_loss.py:282-285)
//   * arg-max (loss/bacs_loss.py:255)
// The gradient rows overwrite the logits in shared memory and leave through TMA bulk
// stores, so HBM sees one read and one write of [B,K,H,W] plus 8+8(+1) bytes per pixel.
#include "common.cuh"

namespace bacs {

struct PixelParams {
  bacs_pixel_args a;
  int P;                // pixels per tile
  int tiles_per_image;
  int use_bulk;         // rows are 16-byte aligned -> TMA bulk copies
  float inv_n;          // 1 / (B*H*W)
  float sy, sx;         // align_corners=True scales (h-1)/(H-1), (w-1)/(W-1)
  double* partials;     // [n_tiles, BACS_NACC]
};

// ---- PTX wrappers: mbarrier + 1-D bulk async copy (TMA) ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- vector access to PPT adjacent pixels of one shared-memory row -------------------
template <typename T, int PPT> struct Vec;
template <> struct Vec<float, 1> {
  __device__ static __forceinline__ void ld(const float* p, float* v) { v[0] = p[0]; }
  __device__ static __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
};
template <> struct Vec<float, 2> {
  __device__ static __forceinline__ void ld(const float* p, float* v) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
  __device__ static __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
};
template <> struct Vec<__nv_bfloat16, 1> {
  __device__ static __forceinline__ void ld(const __nv_bfloat16* p, float* v) { v[0] = __bfloat162float(p[0]); }
  __device__ static __forceinline__ void st(__nv_bfloat16* p, const float* v) { p[0] = __float2bfloat16_rn(v[0]); }
};
template <> struct Vec<__nv_bfloat16, 2> {
  __device__ static __forceinline__ void ld(const __nv_bfloat16* p, float* v) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    v[0] = t.x; v[1] = t.y;
  }
  __device__ static __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
  }
};
template <> struct Vec<__half, 1> {
  __device__ static __forceinline__ void ld(const __half* p, float* v) { v[0] = __half2float(p[0]); }
  __device__ static __forceinline__ void st(__half* p, const float* v) { p[0] = __float2half_rn(v[0]); }
};
template <> struct Vec<__half, 2> {
  __device__ static __forceinline__ void ld(const __half* p, float* v) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(p));
    v[0] = t.x; v[1] = t.y;
  }
  __device__ static __forceinline__ void st(__half* p, const float* v) {
    *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v[0], v[1]);
  }
};

__device__ __forceinline__ float pow_gamma(float base, float gamma) {
  if (gamma == 2.f) return base * base;
  if (gamma == 1.f) return base;
  if (gamma == 0.f) return 1.f;
  return powf(base, gamma);
}

constexpr float kLog2e = 1.4426950408889634f;

template <typename T, int PPT>
__global__ void __launch_bounds__(256) pixel_loss_kernel(const PixelParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ float red_scratch[32];
  __shared__ float s_norm;  // CE-type gradient normaliser (1 / sum of weights)

  const bacs_pixel_args& a = p.a;
  T* tile = reinterpret_cast<T*>(smem_raw);
  const int P = p.P;
  const int K = a.K;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / p.tiles_per_image;
  const int tile_in_img = blockIdx.x - b * p.tiles_per_image;
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t p0 = (int64_t)tile_in_img * P;
  const int npx = (int)min((int64_t)P, HW - p0);
  const T* __restrict__ src = reinterpret_cast<const T*>(a.logits) + (int64_t)b * K * HW + p0;
  const bool bulk = p.use_bulk && ((npx * (int)sizeof(T)) & 15) == 0;

  // ---- 1. start the tile load -------------------------------------------------------
  if (bulk) {
    if (tid == 0) {
      mbar_init(&mbar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid < 32) {
      if (tid == 0) mbar_expect_tx(&mbar, (uint32_t)(K * npx * (int)sizeof(T)));
      __syncwarp();
      for (int c = tid; c < K; c += 32) bulk_g2s(tile + (int64_t)c * P, src + (int64_t)c * HW, (uint32_t)(npx * sizeof(T)), &mbar);
    }
  } else {
    for (int c = 0; c < K; ++c)
      for (int i = tid; i < npx; i += blockDim.x) tile[(int64_t)c * P + i] = src[(int64_t)c * HW + i];
  }

  // ---- 2. per-pixel side inputs while the copies fly --------------------------------
  const int px0 = tid * PPT;
  int y[PPT];         // label, -1 = ignore / invalid
  bool is_ign[PPT];
  float zmax[PPT], zfoc[PPT];
  int cell[PPT];      // low-res cell id y0*w+x0 (focal scatter key)
  int cell_dx[PPT], cell_dy[PPT];
  float wy1[PPT], wx1[PPT];
  float acc[BACS_NACC];
#pragma unroll
  for (int i = 0; i < BACS_NACC; ++i) acc[i] = 0.f;

  const int64_t* lab = a.labels + (int64_t)b * HW + p0;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int px = px0 + j;
    y[j] = -1;
    is_ign[j] = true;
    zmax[j] = 0.f;
    zfoc[j] = 0.f;
    cell[j] = -1;
    cell_dx[j] = cell_dy[j] = 0;
    wy1[j] = wx1[j] = 0.f;
    if (px < npx) {
      const int64_t l = __ldg(lab + px);
      if (l == a.ignore_index) {
        // ignored
      } else if (l >= 0 && l < K) {
        y[j] = (int)l;
        is_ign[j] = false;
      } else {
        acc[BACS_ACC_INVALID] += 1.f;
      }
      if (a.z) {
        const int64_t pix = p0 + px;
        const int Y = (int)(pix / a.W), X = (int)(pix - (int64_t)Y * a.W);
        const Lerp ly = lerp_align_corners(Y, a.h, p.sy), lx = lerp_align_corners(X, a.w, p.sx);
        const float wx0 = 1.f - lx.w1, wy0 = 1.f - ly.w1;
        const float* zb = a.z + (int64_t)b * a.T * a.h * a.w;
        const int o00 = ly.i0 * a.w + lx.i0, o01 = ly.i0 * a.w + lx.i1;
        const int o10 = ly.i1 * a.w + lx.i0, o11 = ly.i1 * a.w + lx.i1;
        float m = -INFINITY;
        for (int t = 0; t < a.T; ++t) {
          const float* zt = zb + (int64_t)t * a.h * a.w;
          // operation order of ATen's upsample_bilinear2d (no FMA contraction)
          const float top = __fadd_rn(__fmul_rn(wx0, __ldg(zt + o00)), __fmul_rn(lx.w1, __ldg(zt + o01)));
          const float bot = __fadd_rn(__fmul_rn(wx0, __ldg(zt + o10)), __fmul_rn(lx.w1, __ldg(zt + o11)));
          const float v = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(ly.w1, bot));
          m = fmaxf(m, v);
          if (t == a.focal_head) zfoc[j] = v;
        }
        zmax[j] = m;
        cell[j] = o00;
        cell_dx[j] = lx.i1 - lx.i0;
        cell_dy[j] = (ly.i1 - ly.i0) * a.w;
        wy1[j] = ly.w1;
        wx1[j] = lx.w1;
      }
    }
  }

  // CE-type modes: gradient normaliser from the label histogram (device-side, no sync)
  if (a.mode != BACS_PIX_WEIGHTED_CE && a.dlogits != nullptr) {
    if (tid < 32) {
      double s = 0.0;
      if (a.mode == BACS_PIX_CE) {
        for (int c = tid; c < K && c < 256; c += 32)
          if (c != a.ignore_index) s += (double)a.hist[c] * (a.class_w ? (double)a.class_w[c] : 1.0);
      } else {  // UNBIASED_CE: number of non-ignored pixels
        for (int c = tid; c < K && c < 256; c += 32)
          if (c != a.ignore_index) s += (double)a.hist[c];
      }
      s = warp_sum(s);
      if (tid == 0) s_norm = s > 0.0 ? (float)(1.0 / s) : 0.f;
    }
  }

  // ---- 3. wait for the tile ----------------------------------------------------------
  if (bulk) mbar_wait(&mbar, 0);
  __syncthreads();

  // ---- 4. softmax statistics ----------------------------------------------------------
  const bool live = px0 < npx;  // P and npx are multiples of PPT on the vector path
  const int old_cl = min(max(a.old_cl, 0), K);
  float mx[PPT], s_all[PPT], s_old[PPT], e0[PPT], xy[PPT];
  int amax[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    mx[j] = -INFINITY;
    amax[j] = 0;
    s_all[j] = s_old[j] = e0[j] = 0.f;
    xy[j] = 0.f;
  }
  if (live) {
    const T* col = tile + px0;
    float v[PPT];
    for (int c = 0; c < K; ++c) {
      Vec<T, PPT>::ld(col + (int64_t)c * P, v);
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (v[j] > mx[j]) {
          mx[j] = v[j];
          amax[j] = c;
        }
    }
    float nm[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) nm[j] = -mx[j] * kLog2e;
    for (int c = 0; c < old_cl; ++c) {
      Vec<T, PPT>::ld(col + (int64_t)c * P, v);
#pragma unroll
      for (int j = 0; j < PPT; ++j) s_old[j] += exp2f(fmaf(v[j], kLog2e, nm[j]));
    }
#pragma unroll
    for (int j = 0; j < PPT; ++j) s_all[j] = s_old[j];
    for (int c = old_cl; c < K; ++c) {
      Vec<T, PPT>::ld(col + (int64_t)c * P, v);
#pragma unroll
      for (int j = 0; j < PPT; ++j) s_all[j] += exp2f(fmaf(v[j], kLog2e, nm[j]));
    }
    Vec<T, PPT>::ld(col, v);
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      e0[j] = exp2f(fmaf(v[j], kLog2e, nm[j]));
      if (y[j] >= 0) xy[j] = DT<T>::to_f(col[(int64_t)y[j] * P + j]);
    }
  }

  // ---- 5. per-pixel losses and gradient coefficients ----------------------------------
  // gradient of pixel:  g_k = e_k * cg[group(k)] - [k==0]*d0 - [k==y]*dy
  // groups: 0 -> k==0 ; 1 -> 1<=k<old_cl ; 2 -> k>=old_cl
  float cg0[PPT], cg1[PPT], cg2[PPT], d0[PPT], dy[PPT];
  float gfoc[PPT];  // d(focal term)/dZ of the pixel
  uint8_t dmask[PPT];
  const float gs_bacs = p.inv_n * a.grad_scale;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    cg0[j] = cg1[j] = cg2[j] = d0[j] = dy[j] = 0.f;
    gfoc[j] = 0.f;
    dmask[j] = 0;
    if (!(live && px0 + j < npx)) continue;
    const float S = s_all[j];
    const float logS = logf(S);
    const float lse = mx[j] + logS;
    const float inv_S = 1.f / S;
    const bool valid = y[j] >= 0;
    if (!is_ign[j]) acc[BACS_ACC_KEPT] += 1.f;
    if (valid) acc[BACS_ACC_VALID] += 1.f;
    if (valid && y[j] == 0) acc[BACS_ACC_BG] += 1.f;
    float seen = 0.f;
    if (a.seen_max) seen = __ldg(a.seen_max + (int64_t)b * HW + p0 + px0 + j);
    else if (a.z) seen = sigmoid_acc(zmax[j]);

    if (a.mode == BACS_PIX_WEIGHTED_CE) {
      if (valid) {
        const float S_fg = S - e0[j];
        const float u = a.ukd ? 1.f : 0.f;
        const float inv_old = 1.f / s_old[j];
        const float inv_fg = 1.f / S_fg;
        float l1, l2;
        if (y[j] == 0) {
          float s = seen;
          if (s > a.threshold) s = 1.f;
          const float mod = pow_gamma(1.f - s, a.gamma);
          l1 = mod * (lse - DT<T>::to_f(tile[px0 + j]));
          l2 = u * (logS - logf(s_old[j]));
          cg0[j] = mod * inv_S + u * (inv_S - inv_old);
          cg1[j] = cg0[j];
          cg2[j] = mod * inv_S + u * inv_S;
          d0[j] = mod;
        } else if (y[j] < old_cl) {
          l1 = logS - logf(S_fg);
          l2 = u * (logS - logf(s_old[j]));
          cg0[j] = inv_S + u * (inv_S - inv_old);
          cg1[j] = inv_S - inv_fg + u * (inv_S - inv_old);
          cg2[j] = inv_S - inv_fg + u * inv_S;
        } else {
          l1 = logS - logf(S_fg);
          l2 = lse - xy[j];
          cg0[j] = 2.f * inv_S;
          cg1[j] = 2.f * inv_S - inv_fg;
          cg2[j] = cg1[j];
          dy[j] = 1.f;
        }
        acc[BACS_ACC_LOSS] += l1 + l2;
        cg0[j] *= gs_bacs; cg1[j] *= gs_bacs; cg2[j] *= gs_bacs; d0[j] *= gs_bacs; dy[j] *= gs_bacs;
      }
    } else if (a.mode == BACS_PIX_CE || a.mode == BACS_PIX_SCORE) {
      if (valid) {
        const float wgt = a.class_w ? __ldg(a.class_w + y[j]) : 1.f;
        acc[BACS_ACC_LOSS] += wgt * (lse - xy[j]);
        acc[BACS_ACC_WSUM] += wgt;
        if (a.dlogits) {
          const float g = wgt * s_norm * a.grad_scale;
          cg0[j] = cg1[j] = cg2[j] = g * inv_S;
          dy[j] = g;
        }
      }
    } else {  // BACS_PIX_UNBIASED_CE
      if (valid) {
        const float g = a.dlogits ? s_norm * a.grad_scale : 0.f;
        if (y[j] < old_cl) {
          acc[BACS_ACC_LOSS] += logS - logf(s_old[j]);
          cg0[j] = cg1[j] = g * (inv_S - 1.f / s_old[j]);
          cg2[j] = g * inv_S;
        } else {
          acc[BACS_ACC_LOSS] += lse - xy[j];
          cg0[j] = cg1[j] = cg2[j] = g * inv_S;
          dy[j] = g;
        }
        acc[BACS_ACC_WSUM] += 1.f;
      }
    }

    // teacher-distill pixel mask: background label and confidently "seen"
    if (a.distill_mask) {
      const bool m = valid && y[j] == 0 && ((a.z == nullptr && a.seen_max == nullptr) || seen > a.lkd_threshold);
      dmask[j] = m ? 1 : 0;
      if (m) acc[BACS_ACC_DISTILL_PIX] += 1.f;
    }

    // seen-detector focal loss of head `focal_head` (binary, target = foreground)
    if (a.gz && !is_ign[j]) {
      const float Z = zfoc[j];
      const float t = (valid && y[j] == 0) ? 0.f : 1.f;
      const float bce = fmaxf(Z, 0.f) - Z * t + log1pf(expf(-fabsf(Z)));
      const float pt = expf(-bce);
      const float om = 1.f - pt;
      float term = pow_gamma(om, a.focal_gamma) * bce;
      const float sig = sigmoid_acc(Z);
      float dterm;
      if (a.focal_gamma == 2.f) dterm = (sig - t) * (om * om + 2.f * om * pt * bce);
      else if (a.focal_gamma == 0.f) dterm = (sig - t);
      else dterm = (sig - t) * (pow_gamma(om, a.focal_gamma) + a.focal_gamma * powf(om, a.focal_gamma - 1.f) * pt * bce);
      if (a.focal_alpha >= 0.f) {
        const float aw = a.focal_alpha * t + (1.f - a.focal_alpha) * (1.f - t);
        term *= aw;
        dterm *= aw;
      }
      acc[BACS_ACC_FOCAL] += term;
      gfoc[j] = dterm;
    }
  }

  // ---- 6. gradient rows overwrite the tile; arg-max / mask stores ----------------------
  if (a.dlogits && live) {
    T* col = tile + px0;
    float v[PPT], g[PPT], nm[PPT], ey[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      nm[j] = -mx[j] * kLog2e;
      ey[j] = exp2f(fmaf(xy[j], kLog2e, nm[j]));
    }
#pragma unroll
    for (int j = 0; j < PPT; ++j) g[j] = e0[j] * cg0[j] - d0[j];
    Vec<T, PPT>::st(col, g);
    for (int c = 1; c < old_cl; ++c) {
      Vec<T, PPT>::ld(col + (int64_t)c * P, v);
#pragma unroll
      for (int j = 0; j < PPT; ++j) g[j] = exp2f(fmaf(v[j], kLog2e, nm[j])) * cg1[j];
      Vec<T, PPT>::st(col + (int64_t)c * P, g);
    }
    for (int c = max(old_cl, 1); c < K; ++c) {
      Vec<T, PPT>::ld(col + (int64_t)c * P, v);
#pragma unroll
      for (int j = 0; j < PPT; ++j) g[j] = exp2f(fmaf(v[j], kLog2e, nm[j])) * cg2[j];
      Vec<T, PPT>::st(col + (int64_t)c * P, g);
    }
    // the label's own channel: recomputed in fp32 from the saved x_y so the -dy term is
    // applied before rounding to the storage type
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (y[j] >= 0 && dy[j] != 0.f) {
        const int k = y[j];
        const float cgk = k == 0 ? cg0[j] : (k < old_cl ? cg1[j] : cg2[j]);
        const float gv = ey[j] * cgk - dy[j] - (k == 0 ? d0[j] : 0.f);
        col[(int64_t)k * P + j] = DT<T>::from_f(gv);
      }
    }
  }
  if (live) {
    if (a.preds) {
      int64_t* out = a.preds + (int64_t)b * HW + p0 + px0;
      if (PPT == 2 && px0 + 1 < npx && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        *reinterpret_cast<longlong2*>(out) = make_longlong2((long long)amax[0], (long long)amax[PPT - 1]);
      } else {
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (px0 + j < npx) out[j] = amax[j];
      }
    }
    if (a.distill_mask) {
      uint8_t* out = a.distill_mask + (int64_t)b * HW + p0 + px0;
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (px0 + j < npx) out[j] = dmask[j];
    }
  }

  // ---- 7. focal gradient: adjoint of the bilinear up-sample ----------------------------
  // lanes of a warp hold consecutive pixels, so equal low-res cells form runs: segmented
  // warp reduction keyed by the cell id, then one global atomic per run and tap.
  if (a.gz) {
    const unsigned full = 0xffffffffu;
    const int lane = tid & 31;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      // process pixel j of every lane; with PPT == 2 runs are still contiguous per j
      float c00 = gfoc[j] * (1.f - wy1[j]) * (1.f - wx1[j]);
      float c01 = gfoc[j] * (1.f - wy1[j]) * wx1[j];
      float c10 = gfoc[j] * wy1[j] * (1.f - wx1[j]);
      float c11 = gfoc[j] * wy1[j] * wx1[j];
      const int key = cell[j] * 4 + cell_dx[j] + 2 * (cell_dy[j] != 0);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n00 = __shfl_down_sync(full, c00, o), n01 = __shfl_down_sync(full, c01, o);
        const float n10 = __shfl_down_sync(full, c10, o), n11 = __shfl_down_sync(full, c11, o);
        const int nk = __shfl_down_sync(full, key, o);
        if (lane + o < 32 && nk == key) {
          c00 += n00; c01 += n01; c10 += n10; c11 += n11;
        }
      }
      const int pk = __shfl_up_sync(full, key, 1);
      const bool head = (lane == 0) || (pk != key);
      if (head && cell[j] >= 0) {
        float* g = a.gz + (int64_t)b * a.h * a.w + cell[j];
        if (c00 != 0.f) atomicAdd(g, c00);
        if (c01 != 0.f) atomicAdd(g + cell_dx[j], c01);
        if (c10 != 0.f) atomicAdd(g + cell_dy[j], c10);
        if (c11 != 0.f) atomicAdd(g + cell_dy[j] + cell_dx[j], c11);
      }
    }
  }

  // ---- 8. ship the gradient tile ---------------------------------------------------------
  if (a.dlogits) {
    T* dst = reinterpret_cast<T*>(a.dlogits) + (int64_t)b * K * HW + p0;
    if (bulk) {
      fence_proxy_async();
      __syncthreads();
      if (tid < 32) {
        for (int c = tid; c < K; c += 32) bulk_s2g(dst + (int64_t)c * HW, tile + (int64_t)c * P, (uint32_t)(npx * sizeof(T)));
        bulk_commit();
      }
    } else {
      __syncthreads();
      for (int c = 0; c < K; ++c)
        for (int i = tid; i < npx; i += blockDim.x) dst[(int64_t)c * HW + i] = tile[(int64_t)c * P + i];
    }
  }

  // ---- 9. per-tile partial sums ------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < BACS_NACC; ++i) {
    const float r = block_sum(acc[i], red_scratch);
    if (tid == 0) p.partials[(int64_t)blockIdx.x * BACS_NACC + i] = (double)r;
  }
  if (a.dlogits && bulk && tid < 32) bulk_wait_read0();
}

// Deterministic reduction of the per-tile partials: grid.x = 1 (whole batch -> acc) plus,
// in SCORE mode, one block per image (-> score[b] = -sum / (H*W)).
__global__ void __launch_bounds__(1024) pixel_reduce_kernel(const double* __restrict__ partials, int n_tiles,
                                                            int tiles_per_image, double* __restrict__ acc,
                                                            double* __restrict__ score, double inv_hw) {
  __shared__ double scratch[32];
  const bool whole = blockIdx.x == 0;
  const int t0 = whole ? 0 : (blockIdx.x - 1) * tiles_per_image;
  const int t1 = whole ? n_tiles : t0 + tiles_per_image;
  const int nacc = whole ? BACS_NACC : 1;
  for (int i = 0; i < nacc; ++i) {
    double s = 0.0;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) s += partials[(int64_t)t * BACS_NACC + i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
      if (whole) acc[i] = s;
      else score[blockIdx.x - 1] = -s * inv_hw;
    }
  }
}

struct PixelPlan {
  int ppt, threads, P;
  size_t smem;
};

static bool make_plan(const bacs_pixel_args& a, PixelPlan* plan) {
  const size_t es = dtype_size(a.dtype);
  const int64_t HW = (int64_t)a.H * a.W;
  const int cand[4][2] = {{2, 256}, {2, 128}, {1, 128}, {1, 64}};
  const size_t soft = 57 * 1024, hard = 200 * 1024;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < 4; ++i) {
      if (cand[i][0] == 2 && (HW & 1)) continue;  // pixel pairs need even image sizes
      const int P = cand[i][0] * cand[i][1];
      const size_t smem = (size_t)a.K * P * es;
      if (smem <= (pass == 0 ? soft : hard)) {
        plan->ppt = cand[i][0];
        plan->threads = cand[i][1];
        plan->P = P;
        plan->smem = smem;
        return true;
      }
    }
  return false;
}

}  // namespace bacs

using namespace bacs;

extern "C" {

size_t bacs_pixel_workspace_bytes(const bacs_pixel_args* a) {
  if (!a) return 0;
  PixelPlan plan;
  if (!make_plan(*a, &plan)) return 0;
  const int64_t HW = (int64_t)a->H * a->W;
  const int64_t tiles = (HW + plan.P - 1) / plan.P * a->B;
  return align_up((size_t)tiles * BACS_NACC * sizeof(double), 256);
}

int bacs_pixel_loss(const bacs_pixel_args* a, void* workspace, size_t workspace_bytes, bacs_stream_t stream) {
  BACS_REQUIRE(a, "bacs_pixel_loss: null args");
  BACS_REQUIRE(a->logits && a->labels && a->acc, "bacs_pixel_loss: logits, labels and acc are required");
  BACS_REQUIRE(a->B > 0 && a->K > 0 && a->H > 0 && a->W > 0, "bacs_pixel_loss: bad shape");
  BACS_REQUIRE(a->K <= 255 || a->ignore_index >= a->K || a->ignore_index < 0,
               "bacs_pixel_loss: ignore_index inside the class range");
  BACS_REQUIRE(a->mode >= 0 && a->mode <= BACS_PIX_SCORE, "bacs_pixel_loss: unknown mode %d", a->mode);
  BACS_REQUIRE(a->dtype >= 0 && a->dtype <= BACS_F16, "bacs_pixel_loss: unknown dtype %d", a->dtype);
  if (a->z) {
    BACS_REQUIRE(a->T > 0 && a->h > 0 && a->w > 0, "bacs_pixel_loss: seen logits given without T/h/w");
    BACS_REQUIRE(a->seen_scale > 0 && a->H == a->h * a->seen_scale && a->W == a->w * a->seen_scale,
                 "bacs_pixel_loss: H,W must equal h,w * seen_scale (the reference's nn.Upsample(scale_factor))");
  }
  if (a->gz) BACS_REQUIRE(a->z && a->focal_head >= 0 && a->focal_head < a->T, "bacs_pixel_loss: focal head out of range");
  if (a->mode == BACS_PIX_WEIGHTED_CE)
    BACS_REQUIRE(a->old_cl >= 1 && (a->z || a->seen_max),
                 "bacs_pixel_loss: WEIGHTED_CE needs old_cl >= 1 and the seen logits / probabilities");
  if (a->mode != BACS_PIX_WEIGHTED_CE && a->dlogits)
    BACS_REQUIRE(a->hist, "bacs_pixel_loss: CE-type gradients need the label histogram");
  if (a->mode == BACS_PIX_SCORE) BACS_REQUIRE(a->score && !a->dlogits, "bacs_pixel_loss: SCORE mode needs score and no gradient");
  PixelPlan plan;
  if (!make_plan(*a, &plan)) {
    set_error("bacs_pixel_loss: K=%d too large for a shared-memory tile", a->K);
    return BACS_ERR_UNSUPPORTED;
  }
  const int64_t HW = (int64_t)a->H * a->W;
  const int64_t tiles_per_image = (HW + plan.P - 1) / plan.P;
  const int64_t n_tiles = tiles_per_image * a->B;
  BACS_REQUIRE(n_tiles < 0x7fffffff, "bacs_pixel_loss: too many tiles");
  if (workspace_bytes < (size_t)n_tiles * BACS_NACC * sizeof(double) || !workspace) {
    set_error("bacs_pixel_loss: workspace too small (%zu bytes)", workspace_bytes);
    return BACS_ERR_WORKSPACE;
  }
  const size_t es = dtype_size(a->dtype);
  PixelParams p;
  p.a = *a;
  p.P = plan.P;
  p.tiles_per_image = (int)tiles_per_image;
  p.inv_n = (float)(1.0 / ((double)a->B * (double)HW));
  p.sy = a->z ? ac_scale(a->h, a->H) : 0.f;
  p.sx = a->z ? ac_scale(a->w, a->W) : 0.f;
  p.partials = reinterpret_cast<double*>(workspace);
  const bool aligned = ((HW * es) % 16 == 0) && ((reinterpret_cast<uintptr_t>(a->logits) & 15) == 0) &&
                       (!a->dlogits || (reinterpret_cast<uintptr_t>(a->dlogits) & 15) == 0);
  p.use_bulk = aligned ? 1 : 0;
  const int64_t tiles2 = n_tiles;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_PIX(TT, PPT)                                                                                  \
  do {                                                                                                       \
    auto kern = pixel_loss_kernel<TT, PPT>;                                                                  \
    if (plan.smem > 48 * 1024) {                                                                             \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem); \
      if (e != cudaSuccess) {                                                                                \
        set_error("bacs_pixel_loss: cannot opt in to %zu bytes of shared memory: %s", plan.smem,            \
                  cudaGetErrorString(e));                                                                    \
        return BACS_ERR_CUDA;                                                                                \
      }                                                                                                      \
    }                                                                                                        \
    kern<<<(unsigned)tiles2, plan.threads, plan.smem, s>>>(p);                                               \
  } while (0)
  BACS_DISPATCH_DTYPE(a->dtype, TT, {
    if (plan.ppt == 2) LAUNCH_PIX(TT, 2);
    else LAUNCH_PIX(TT, 1);
  });
#undef LAUNCH_PIX
  BACS_CHECK_LAUNCH("bacs_pixel_loss");
  const int nblk = 1 + (a->mode == BACS_PIX_SCORE ? a->B : 0);
  pixel_reduce_kernel<<<nblk, 1024, 0, s>>>(p.partials, (int)tiles2, p.tiles_per_image, a->acc, a->score,
                                            1.0 / (double)HW);
  BACS_CHECK_LAUNCH("bacs_pixel_loss(reduce)");
  return BACS_OK;
}

}  // extern "C"
