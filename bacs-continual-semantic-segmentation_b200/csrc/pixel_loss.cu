// The fused per-pixel kernel of the BACS loss step.
//
// Persistent, warp-specialised CTAs (one producer warp + 8 consumer warps).  A tile is P
// consecutive pixels of one image; its K logit rows ([K][P], NCHW so every row is one
// contiguous segment) travel global -> shared memory by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier) through a ring of stages.  Consumers pull the tile into
// registers, compute the softmax statistics ONCE per pixel and derive from them, in the
// same pass:
//   * background-aware unbiased CE  (training/loss_utils.py:542-585)   mode WEIGHTED_CE
//   * plain / class-weighted CE     (loss/base_loss.py:237-240)        mode CE
//   * MiB unbiased CE               (training/loss_utils.py:492-520)   mode UNBIASED_CE
//   * per-image importance score    (loss/bacs_loss.py:183-189)        mode SCORE
//   * seen-detector focal loss of one head + its gradient w.r.t. the low-res head output
//     (loss/base_loss.py:255-272, smp FocalLoss binary; bilinear x16 align_corners=True
//      evaluated on the fly, networks/bg_detector.py:13-15)
//   * the teacher-distill pixel mask (loss/bacs_loss.py:282-285)
//   * arg-max (loss/bacs_loss.py:255)
// The gradient rows overwrite the logits of the stage in place and leave through TMA bulk
// stores issued by the producer warp, so HBM sees one read and one write of [B,K,H,W]
// plus 8 + 8 (+1) bytes per pixel, and the consumers never wait for a store.
#include <algorithm>
#include <cstdlib>

#include "pixel_fast.cuh"  // Raw<T>: packed pixel pairs

namespace bacs {

// Shared-memory layout (dynamic): [stages][K*PITCH] tiles | zr[T][w] | gacc[w+1]
// COOP (large K, one pixel per thread for everything per-pixel): the two threads of an even/odd lane pair walk the
// channel loops together on their two adjacent pixels as PACKED pairs, each taking every other channel, and
// exchange max / arg-max / sums / coefficients by shuffle.  Rows are 64 bytes longer than P elements so that the
// two halves of a warp (rows c and c+1) hit disjoint banks.
template <typename T, int PPT, int KREG, bool ROWTILE, bool COOP = false>
__global__ void __launch_bounds__(kConsumers + 32, ((KREG > 0 || PPT == 1) ? 2 : 1)) pixel_loss_kernel(const PixelParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar_full[kMaxStages];  // producer -> consumers: tile landed
  __shared__ uint64_t bar_done[kMaxStages];  // consumers -> producer: gradients written; store / refill the stage
  __shared__ float red_scratch[kConsumerWarps][BACS_NACC];
  __shared__ float s_norm_sh;

  const bacs_pixel_args& a = p.a;
  const int P = p.P, K = a.K, S = p.stages;
  const int PITCH = COOP ? P + 64 / (int)sizeof(T) : P;  // elements between channel rows of a tile
  const int tid = threadIdx.x;
  const int64_t HW = (int64_t)a.H * a.W;
  const size_t tile_elems = (size_t)K * PITCH;
  T* tiles = reinterpret_cast<T*>(smem_raw);
  float* zr = reinterpret_cast<float*>(smem_raw + ((S * tile_elems * sizeof(T) + 15) / 16) * 16);
  float* gacc = zr + (ROWTILE ? a.T * a.w : 0);
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_done[s], kConsumers);
    }
    fence_mbar_init();
    s_norm_sh = 0.f;
  }
  __syncthreads();
  // CE-type modes: gradient normaliser from the label histogram (device-side, no host sync)
  if (tid < 32 && a.mode != BACS_PIX_WEIGHTED_CE && a.dlogits != nullptr) {
    double s = 0.0;
    for (int c = tid; c < K && c < 256; c += 32)
      if (c != a.ignore_index)
        s += (double)a.hist[c] * ((a.mode == BACS_PIX_CE && a.class_w) ? (double)a.class_w[c] : 1.0);
    s = warp_sum(s);
    if (tid == 0) s_norm_sh = s > 0.0 ? (float)(1.0 / s) : 0.f;
  }
  __syncthreads();

  // =====================================================================================
  // producer warp
  // =====================================================================================
  if (tid >= kConsumers) {
    const int lane = tid - kConsumers;
    auto tile_geom = [&](int k, int& b, int64_t& p0, int& npx) {
      const int tile = (int)blockIdx.x + k * (int)gridDim.x;
      b = tile / p.tiles_per_image;
      p0 = (int64_t)(tile - b * p.tiles_per_image) * P;
      npx = (int)min((int64_t)P, HW - p0);
    };
    auto load_tile = [&](int k) {
      int b, npx;
      int64_t p0;
      tile_geom(k, b, p0, npx);
      const int s = k % S;
      T* dst = tiles + (size_t)s * tile_elems;
      const T* src = reinterpret_cast<const T*>(a.logits) + (int64_t)b * K * HW + p0;
      const bool bulk = p.use_bulk && ((npx * (int)sizeof(T)) & 15) == 0;
      if (bulk) {
        if (lane == 0) mbar_expect_tx(&bar_full[s], (uint32_t)(K * npx * (int)sizeof(T)));
        __syncwarp();
        for (int c = lane; c < K; c += 32)
          bulk_g2s(dst + (size_t)c * PITCH, src + (int64_t)c * HW, (uint32_t)(npx * sizeof(T)), &bar_full[s]);
      } else {  // ragged / unaligned rows: plain copies by the producer warp
        for (int c = 0; c < K; ++c)
          for (int i = lane; i < npx; i += 32) dst[(size_t)c * PITCH + i] = src[(int64_t)c * HW + i];
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[s]);
      }
    };
    const int pre = min(S, my_tiles);
    for (int k = 0; k < pre; ++k) load_tile(k);
    for (int k = 0; k < my_tiles; ++k) {
      const int s = k % S;
      mbar_wait(&bar_done[s], (uint32_t)((k / S) & 1));
      if (a.dlogits) {
        int b, npx;
        int64_t p0;
        tile_geom(k, b, p0, npx);
        const T* src = tiles + (size_t)s * tile_elems;
        T* dst = reinterpret_cast<T*>(a.dlogits) + (int64_t)b * K * HW + p0;
        const bool bulk = p.use_bulk && ((npx * (int)sizeof(T)) & 15) == 0;
        if (bulk) {
          for (int c = lane; c < K; c += 32)
            bulk_s2g(dst + (int64_t)c * HW, src + (size_t)c * PITCH, (uint32_t)(npx * sizeof(T)));
          bulk_commit();
          if (k + S < my_tiles) bulk_wait_read0();  // the stage is about to be refilled
        } else {
          for (int c = 0; c < K; ++c)
            for (int i = lane; i < npx; i += 32) dst[(int64_t)c * HW + i] = src[(size_t)c * PITCH + i];
        }
        __syncwarp();
      }
      if (k + S < my_tiles) load_tile(k + S);
    }
    bulk_wait_all();
    return;
  }

  // =====================================================================================
  // consumer warps
  // =====================================================================================
  const int lane = tid & 31, wid = tid >> 5;
  const int px0 = tid * PPT;
  const int old_cl = min(max(a.old_cl, 0), K);
  const bool have_seen = (a.z != nullptr) || (a.seen_max != nullptr);
  const float s_norm = s_norm_sh;
  float acc[BACS_NACC];
#pragma unroll
  for (int i = 0; i < BACS_NACC; ++i) acc[i] = 0.f;

  auto tile_geom = [&](int k, int& b, int64_t& p0, int& npx) {
    const int tile = (int)blockIdx.x + k * (int)gridDim.x;
    b = tile / p.tiles_per_image;
    p0 = (int64_t)(tile - b * p.tiles_per_image) * P;
    npx = (int)min((int64_t)P, HW - p0);
  };
  auto load_labels = [&](int k, int64_t* lab) {
    int b, npx;
    int64_t p0;
    tile_geom(k, b, p0, npx);
    const int64_t* src = a.labels + (int64_t)b * HW + p0 + px0;
    if (PPT == 2 && px0 + 1 < npx && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(src));
      lab[0] = v.x;
      lab[PPT - 1] = v.y;
    } else {
#pragma unroll
      for (int j = 0; j < PPT; ++j) lab[j] = (px0 + j < npx) ? __ldg(src + j) : (int64_t)a.ignore_index;
    }
  };
  int64_t lab_next[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) lab_next[j] = a.ignore_index;
  if (my_tiles > 0) load_labels(0, lab_next);

  for (int k = 0; k < my_tiles; ++k) {
    int b, npx;
    int64_t p0;
    tile_geom(k, b, p0, npx);
    const int s = k % S;
    T* tile = tiles + (size_t)s * tile_elems;
    int64_t lab[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) lab[j] = lab_next[j];
    if (k + 1 < my_tiles) load_labels(k + 1, lab_next);

    // ---- per-pixel side inputs --------------------------------------------------------
    int y[PPT];
    bool is_ign[PPT];
    float seen[PPT], zfoc[PPT], wx1[PPT];
    int cx0[PPT], cdx[PPT];
    int cell[PPT], cell_dy[PPT];  // generic (non row-tile) focal scatter bookkeeping
    float wy1g[PPT];
    int Yrow = 0;
    Lerp ly_row = {0, 0, 0.f};
    if (ROWTILE && a.z) {
      // the tile lies inside image row Yrow: interpolate the T head rows in y once
      Yrow = (int)(p0 / a.W);
      ly_row = lerp_align_corners(Yrow, a.h, p.sy);
      consumer_sync();  // previous tile's readers of zr / gacc are done
      const float* zb = a.z + (int64_t)b * a.T * a.h * a.w;
      const float wy0 = 1.f - ly_row.w1;
      for (int i = tid; i < a.T * a.w; i += kConsumers) {
        const int t = i / a.w, j = i - t * a.w;
        const float* zt = zb + (int64_t)t * a.h * a.w;
        zr[i] = __fadd_rn(__fmul_rn(wy0, __ldg(zt + ly_row.i0 * a.w + j)),
                          __fmul_rn(ly_row.w1, __ldg(zt + ly_row.i1 * a.w + j)));
      }
      if (a.gz)
        for (int i = tid; i < a.w + 1; i += kConsumers) gacc[i] = 0.f;
      consumer_sync();
    }
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int px = px0 + j;
      y[j] = -1;
      is_ign[j] = true;
      seen[j] = 0.f;
      zfoc[j] = 0.f;
      wx1[j] = 0.f;
      cx0[j] = -1;
      cdx[j] = 0;
      cell[j] = -1;
      cell_dy[j] = 0;
      wy1g[j] = 0.f;
      if (px < npx) {
        const int64_t l = lab[j];
        if (l == a.ignore_index) {
        } else if (l >= 0 && l < K) {
          y[j] = (int)l;
          is_ign[j] = false;
        } else {
          acc[BACS_ACC_INVALID] += 1.f;
        }
        if (a.seen_max) seen[j] = __ldg(a.seen_max + (int64_t)b * HW + p0 + px);
        if (a.z) {
          const int64_t pix = p0 + px;
          float zmax = -INFINITY;
          if (ROWTILE) {
            const int X = (int)(pix - (int64_t)Yrow * a.W);
            const Lerp lx = lerp_align_corners(X, a.w, p.sx);
            const float wx0 = 1.f - lx.w1;
            for (int t = 0; t < a.T; ++t) {
              const float v = __fadd_rn(__fmul_rn(wx0, zr[t * a.w + lx.i0]), __fmul_rn(lx.w1, zr[t * a.w + lx.i1]));
              zmax = fmaxf(zmax, v);
              if (t == a.focal_head) zfoc[j] = v;
            }
            cx0[j] = lx.i0;
            cdx[j] = lx.i1 - lx.i0;
            wx1[j] = lx.w1;
          } else {
            const int Y = (int)(pix / a.W), X = (int)(pix - (int64_t)Y * a.W);
            const Lerp ly = lerp_align_corners(Y, a.h, p.sy), lx = lerp_align_corners(X, a.w, p.sx);
            const float wx0 = 1.f - lx.w1, wy0 = 1.f - ly.w1;
            const float* zb = a.z + (int64_t)b * a.T * a.h * a.w;
            const int o00 = ly.i0 * a.w + lx.i0, o01 = ly.i0 * a.w + lx.i1;
            const int o10 = ly.i1 * a.w + lx.i0, o11 = ly.i1 * a.w + lx.i1;
            for (int t = 0; t < a.T; ++t) {
              const float* zt = zb + (int64_t)t * a.h * a.w;
              const float left = __fadd_rn(__fmul_rn(wy0, __ldg(zt + o00)), __fmul_rn(ly.w1, __ldg(zt + o10)));
              const float right = __fadd_rn(__fmul_rn(wy0, __ldg(zt + o01)), __fmul_rn(ly.w1, __ldg(zt + o11)));
              const float v = __fadd_rn(__fmul_rn(wx0, left), __fmul_rn(lx.w1, right));
              zmax = fmaxf(zmax, v);
              if (t == a.focal_head) zfoc[j] = v;
            }
            cell[j] = o00;
            cdx[j] = lx.i1 - lx.i0;
            cell_dy[j] = (ly.i1 - ly.i0) * a.w;
            wy1g[j] = ly.w1;
            wx1[j] = lx.w1;
          }
          if (!a.seen_max) seen[j] = sigmoid_fast(zmax);
        }
      }
    }

    // ---- wait for the tile, softmax statistics, gradients ------------------------------
    mbar_wait(&bar_full[s], (uint32_t)((k / S) & 1));
    const bool live = px0 < npx;
    const unsigned live_mask = __ballot_sync(0xffffffffu, live);  // (cooperative path: pairs are live together)
    PixCoef pc[PPT];
    float gfoc[PPT];
    uint8_t dmask[PPT];
    int amax[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      gfoc[j] = 0.f;
      dmask[j] = 0;
      amax[j] = 0;
    }
    if (live) {
      T* col = tile + px0;
      float xy[PPT], x0[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        xy[j] = (y[j] >= 0) ? DT<T>::to_f(col[(size_t)y[j] * PITCH + j]) : 0.f;
        x0[j] = DT<T>::to_f(col[j]);
      }
      if (KREG > 0) {
        // ---------------- register-resident path (K <= KREG) ----------------
        constexpr int KR = KREG > 0 ? KREG : 1;
        float e[KR][PPT];
        float mx[PPT], nm[PPT], s_all[PPT], s_old[PPT], s_fg[PPT];
#pragma unroll
        for (int c = 0; c < KR; ++c)
          if (c < K) Vec<T, PPT>::ld(col + (size_t)c * P, e[c]);
#pragma unroll
        for (int j = 0; j < PPT; ++j) mx[j] = e[0][j];
#pragma unroll
        for (int c = 1; c < KR; ++c)
          if (c < K) {
#pragma unroll
            for (int j = 0; j < PPT; ++j)
              if (e[c][j] > mx[j]) {
                mx[j] = e[c][j];
                amax[j] = c;
              }
          }
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          nm[j] = -mx[j] * kLog2e;
          s_fg[j] = s_old[j] = 0.f;
        }
        // the foreground sum is accumulated without channel 0 (S - e_0 would cancel when the background dominates)
#pragma unroll
        for (int c = 0; c < KR; ++c)
          if (c < K) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              e[c][j] = ex2_fast(fmaf(e[c][j], kLog2e, nm[j]));
              if (c >= 1) {
                s_fg[j] += e[c][j];
                if (c < old_cl) s_old[j] += e[c][j];
              }
            }
          }
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          s_all[j] = s_fg[j] + e[0][j];
          if (old_cl >= 1) s_old[j] += e[0][j];
          pixel_terms(a, p.inv_n, s_norm, old_cl, y[j], is_ign[j], mx[j], s_all[j], s_old[j], s_fg[j], e[0][j], x0[j],
                      xy[j], seen[j], have_seen, zfoc[j], acc, pc[j], gfoc[j], dmask[j]);
        }
        if (a.dlogits) {
          float g[PPT];
#pragma unroll
          for (int c = 0; c < KR; ++c)
            if (c < K) {
#pragma unroll
              for (int j = 0; j < PPT; ++j) {
                const float cg = c == 0 ? pc[j].cg0 : (c < old_cl ? pc[j].cg1 : pc[j].cg2);
                g[j] = e[c][j] * cg;
                if (c == 0) g[j] -= pc[j].d0;
                if (c == y[j]) g[j] -= pc[j].dy;
              }
              Vec<T, PPT>::st(col + (size_t)c * PITCH, g);
            }
        }
      } else if constexpr (COOP) {
        // ---------------- cooperative packed path (large K) ----------------
        static_assert(!COOP || PPT == 1, "one pixel per thread");
        const unsigned full = live_mask;
        const int half = tid & 1;                    // this thread takes channels half, half + 2, ...
        T* colp = tile + (px0 & ~1);                 // the pair's first pixel
        // pass 1: packed max / arg-max over my channels, then merged with the partner's (ties -> lowest channel)
        // (channel rows are walked with running pointers: two rows per step)
        const size_t step = 2 * (size_t)PITCH;
        const T* rp = colp + (size_t)half * PITCH;
        typename Raw<T>::Max mt = Raw<T>::init(Raw<T>::ld(rp));
        Raw<T>::set_first(mt, half);
        rp += step;
#pragma unroll 4
        for (int c = half + 2; c < K; c += 2, rp += step) Raw<T>::update(mt, Raw<T>::ld(rp), c);
        float m0, m1;
        int a0, a1;
        Raw<T>::finish(mt, m0, m1, a0, a1);
        {
          const float pm0 = __shfl_xor_sync(full, m0, 1), pm1 = __shfl_xor_sync(full, m1, 1);
          const int pa0 = __shfl_xor_sync(full, a0, 1), pa1 = __shfl_xor_sync(full, a1, 1);
          if (pm0 > m0 || (pm0 == m0 && pa0 < a0)) { m0 = pm0; a0 = pa0; }
          if (pm1 > m1 || (pm1 == m1 && pa1 < a1)) { m1 = pm1; a1 = pa1; }
        }
        // pass 2: exponent sums of both pixels as packed pairs; old and new classes in separate loops
        const F2 nm2 = f2(-m0 * kLog2e, -m1 * kLog2e), l2e2 = f2b(kLog2e);
        auto exp_at = [&](const T* q) {
          float v0, v1;
          Raw<T>::unpack(Raw<T>::ld(q), v0, v1);
          const F2 arg = fma2(f2(v0, v1), l2e2, nm2);
          return f2(ex2_fast(f2lo(arg)), ex2_fast(f2hi(arg)));
        };
        const int new0 = max(old_cl, 1) + ((max(old_cl, 1) ^ half) & 1);  // my first new-class channel (never 0)
        // (channel 0 stays out of the loops: the foreground sum S_fg is accumulated directly, S - e_0 would cancel
        // when the background dominates)
        F2 so2 = f2b(0.f), sn2 = f2b(0.f);
        rp = colp + (size_t)(half == 0 ? 2 : 1) * PITCH;
#pragma unroll 4
        for (int c = (half == 0 ? 2 : 1); c < old_cl; c += 2, rp += step) so2 = add2(so2, exp_at(rp));
        rp = colp + (size_t)new0 * PITCH;
#pragma unroll 4
        for (int c = new0; c < K; c += 2, rp += step) sn2 = add2(sn2, exp_at(rp));
        {
          F2 po, pn;
          po.v = __shfl_xor_sync(full, so2.v, 1);
          pn.v = __shfl_xor_sync(full, sn2.v, 1);
          so2 = add2(so2, po);
          sn2 = add2(sn2, pn);
        }
        const F2 e02 = exp_at(colp);
        const F2 sf2 = add2(so2, sn2);  // foreground channels only
        const F2 sa2 = add2(sf2, e02);
        if (old_cl >= 1) so2 = add2(so2, e02);
        // per-pixel terms: thread `half` owns pixel `half` of the pair; coefficients are swapped by shuffle
        const float mx_me = half ? m1 : m0;
        amax[0] = half ? a1 : a0;
        pixel_terms(a, p.inv_n, s_norm, old_cl, y[0], is_ign[0], mx_me, half ? f2hi(sa2) : f2lo(sa2),
                    half ? f2hi(so2) : f2lo(so2), half ? f2hi(sf2) : f2lo(sf2), half ? f2hi(e02) : f2lo(e02), x0[0], xy[0],
                    seen[0], have_seen, zfoc[0], acc, pc[0], gfoc[0], dmask[0]);
        if (a.dlogits) {
          PixCoef po;
          po.cg0 = __shfl_xor_sync(full, pc[0].cg0, 1);
          po.cg1 = __shfl_xor_sync(full, pc[0].cg1, 1);
          po.cg2 = __shfl_xor_sync(full, pc[0].cg2, 1);
          po.d0 = __shfl_xor_sync(full, pc[0].d0, 1);
          const PixCoef& pa = half ? po : pc[0];   // pixel 0 of the pair
          const PixCoef& pb = half ? pc[0] : po;   // pixel 1 of the pair
          const F2 cgo = f2(pa.cg1, pb.cg1), cgn = f2(pa.cg2, pb.cg2);
          // pass 3: gradient rows of my channels, both pixels at once
          if (half == 0) {
            const F2 g = fma2(e02, f2(pa.cg0, pb.cg0), f2(-pa.d0, -pb.d0));
            Raw<T>::st(colp, f2lo(g), f2hi(g));
          }
          {
            const int c0 = half == 0 ? 2 : 1;
            T* wp = colp + (size_t)c0 * PITCH;
#pragma unroll 4
            for (int c = c0; c < old_cl; c += 2, wp += step) {
              const F2 g = mul2(exp_at(wp), cgo);
              Raw<T>::st(wp, f2lo(g), f2hi(g));
            }
            const int n1 = max(old_cl, 1) + ((max(old_cl, 1) ^ half) & 1);
            wp = colp + (size_t)n1 * PITCH;
#pragma unroll 4
            for (int c = n1; c < K; c += 2, wp += step) {
              const F2 g = mul2(exp_at(wp), cgn);
              Raw<T>::st(wp, f2lo(g), f2hi(g));
            }
          }
          __syncwarp(full);  // the partner's packed row stores are done before single elements are patched
          // the label's own channel, recomputed in fp32 so -dy is applied before rounding
          if (y[0] >= 0 && pc[0].dy != 0.f) {
            const int kk = y[0];
            const float cgk = kk == 0 ? pc[0].cg0 : (kk < old_cl ? pc[0].cg1 : pc[0].cg2);
            const float ey = ex2_fast(fmaf(xy[0], kLog2e, -mx_me * kLog2e));
            col[(size_t)kk * PITCH] = DT<T>::from_f(ey * cgk - pc[0].dy - (kk == 0 ? pc[0].d0 : 0.f));
          }
        }
      } else {
        // ---------------- generic path: three passes over the shared-memory tile ----------------
        float mx[PPT], nm[PPT], s_all[PPT], s_old[PPT], s_fg[PPT], e0[PPT], v[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          mx[j] = -INFINITY;
          s_fg[j] = s_old[j] = 0.f;
        }
        #pragma unroll 8
        for (int c = 0; c < K; ++c) {
          Vec<T, PPT>::ld(col + (size_t)c * PITCH, v);
#pragma unroll
          for (int j = 0; j < PPT; ++j)
            if (v[j] > mx[j]) {
              mx[j] = v[j];
              amax[j] = c;
            }
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j) nm[j] = -mx[j] * kLog2e;
        // channel 0 stays out of the loops: the foreground sum is accumulated directly (S - e_0 would cancel when
        // the background dominates)
        #pragma unroll 8
        for (int c = 1; c < old_cl; ++c) {
          Vec<T, PPT>::ld(col + (size_t)c * PITCH, v);
#pragma unroll
          for (int j = 0; j < PPT; ++j) s_old[j] += ex2_fast(fmaf(v[j], kLog2e, nm[j]));
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j) s_fg[j] = s_old[j];
        #pragma unroll 8
        for (int c = max(old_cl, 1); c < K; ++c) {
          Vec<T, PPT>::ld(col + (size_t)c * PITCH, v);
#pragma unroll
          for (int j = 0; j < PPT; ++j) s_fg[j] += ex2_fast(fmaf(v[j], kLog2e, nm[j]));
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          e0[j] = ex2_fast(fmaf(x0[j], kLog2e, nm[j]));
          s_all[j] = s_fg[j] + e0[j];
          if (old_cl >= 1) s_old[j] += e0[j];
          pixel_terms(a, p.inv_n, s_norm, old_cl, y[j], is_ign[j], mx[j], s_all[j], s_old[j], s_fg[j], e0[j], x0[j],
                      xy[j], seen[j], have_seen, zfoc[j], acc, pc[j], gfoc[j], dmask[j]);
        }
        if (a.dlogits) {
          float g[PPT];
#pragma unroll
          for (int j = 0; j < PPT; ++j) g[j] = e0[j] * pc[j].cg0 - pc[j].d0;
          Vec<T, PPT>::st(col, g);
          #pragma unroll 8
          for (int c = 1; c < old_cl; ++c) {
            Vec<T, PPT>::ld(col + (size_t)c * PITCH, v);
#pragma unroll
            for (int j = 0; j < PPT; ++j) g[j] = ex2_fast(fmaf(v[j], kLog2e, nm[j])) * pc[j].cg1;
            Vec<T, PPT>::st(col + (size_t)c * PITCH, g);
          }
          #pragma unroll 8
          for (int c = max(old_cl, 1); c < K; ++c) {
            Vec<T, PPT>::ld(col + (size_t)c * PITCH, v);
#pragma unroll
            for (int j = 0; j < PPT; ++j) g[j] = ex2_fast(fmaf(v[j], kLog2e, nm[j])) * pc[j].cg2;
            Vec<T, PPT>::st(col + (size_t)c * PITCH, g);
          }
          // the label's own channel, recomputed in fp32 so -dy is applied before rounding
#pragma unroll
          for (int j = 0; j < PPT; ++j) {
            if (y[j] >= 0 && pc[j].dy != 0.f) {
              const int kk = y[j];
              const float cgk = kk == 0 ? pc[j].cg0 : (kk < old_cl ? pc[j].cg1 : pc[j].cg2);
              const float ey = ex2_fast(fmaf(xy[j], kLog2e, nm[j]));
              col[(size_t)kk * PITCH + j] = DT<T>::from_f(ey * cgk - pc[j].dy - (kk == 0 ? pc[j].d0 : 0.f));
            }
          }
        }
      }
    }
    // hand the stage back to the producer (gradient rows are complete)
    fence_proxy_async();
    mbar_arrive(&bar_done[s]);

    // ---- arg-max / mask stores -------------------------------------------------------------
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
        if (PPT == 2 && px0 + 1 < npx) {
          *reinterpret_cast<uchar2*>(out) = make_uchar2(dmask[0], dmask[PPT - 1]);
        } else {
#pragma unroll
          for (int j = 0; j < PPT; ++j)
            if (px0 + j < npx) out[j] = dmask[j];
        }
      }
    }

    // ---- focal gradient: adjoint of the bilinear up-sample ----------------------------------
    if (a.gz) {
      const unsigned full = 0xffffffffu;
      if (ROWTILE) {
        // all pixels share the row weights; reduce (g*(1-wx), g*wx) per low-res column.
        // A thread's PPT pixels are adjacent: combine them when they share the column.
        float c0 = gfoc[0] * (1.f - wx1[0]), c1 = gfoc[0] * wx1[0];
        const int key = cx0[0] * 2 + cdx[0];
        float d0 = 0.f, d1 = 0.f;
        int x2 = -1, dx2 = 0;
        if (PPT == 2) {
          const float f0 = gfoc[PPT - 1] * (1.f - wx1[PPT - 1]), f1 = gfoc[PPT - 1] * wx1[PPT - 1];
          const int k2 = cx0[PPT - 1] * 2 + cdx[PPT - 1];
          if (k2 == key) {
            c0 += f0;
            c1 += f1;
          } else {  // column boundary inside the pair: flushed on its own
            d0 = f0; d1 = f1; x2 = cx0[PPT - 1]; dx2 = cdx[PPT - 1];
          }
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n0 = __shfl_down_sync(full, c0, o), n1 = __shfl_down_sync(full, c1, o);
          const int nk = __shfl_down_sync(full, key, o);
          if (lane + o < 32 && nk == key) {
            c0 += n0;
            c1 += n1;
          }
        }
        const int pk = __shfl_up_sync(full, key, 1);
        if (((lane == 0) || (pk != key)) && cx0[0] >= 0) {
          if (c0 != 0.f) atomicAdd(&gacc[cx0[0]], c0);
          if (c1 != 0.f) atomicAdd(&gacc[cx0[0] + cdx[0]], c1);
        }
        if (x2 >= 0) {
          if (d0 != 0.f) atomicAdd(&gacc[x2], d0);
          if (d1 != 0.f) atomicAdd(&gacc[x2 + dx2], d1);
        }
        consumer_sync();
        float* g = a.gz + (int64_t)b * a.h * a.w;
        const float wy0 = 1.f - ly_row.w1;
        for (int i = tid; i < a.w; i += kConsumers) {
          const float v = gacc[i];
          if (v != 0.f) {
            atomicAdd(g + ly_row.i0 * a.w + i, wy0 * v);
            if (ly_row.w1 != 0.f) atomicAdd(g + ly_row.i1 * a.w + i, ly_row.w1 * v);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          float c00 = gfoc[j] * (1.f - wy1g[j]) * (1.f - wx1[j]);
          float c01 = gfoc[j] * (1.f - wy1g[j]) * wx1[j];
          float c10 = gfoc[j] * wy1g[j] * (1.f - wx1[j]);
          float c11 = gfoc[j] * wy1g[j] * wx1[j];
          const int key = cell[j] * 4 + cdx[j] + 2 * (cell_dy[j] != 0);
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
          if (((lane == 0) || (pk != key)) && cell[j] >= 0) {
            float* g = a.gz + (int64_t)b * a.h * a.w + cell[j];
            if (c00 != 0.f) atomicAdd(g, c00);
            if (c01 != 0.f) atomicAdd(g + cdx[j], c01);
            if (c10 != 0.f) atomicAdd(g + cell_dy[j], c10);
            if (c11 != 0.f) atomicAdd(g + cell_dy[j] + cdx[j], c11);
          }
        }
      }
    }

    // SCORE mode needs per-image sums: flush the accumulators per tile
    if (a.mode == BACS_PIX_SCORE) {
#pragma unroll
      for (int i = 0; i < BACS_NACC; ++i) {
        const float r = warp_sum(acc[i]);
        if (lane == 0) red_scratch[wid][i] = r;
        acc[i] = 0.f;
      }
      consumer_sync();
      if (tid < BACS_NACC) {
        float r = 0.f;
        for (int wv = 0; wv < kConsumerWarps; ++wv) r += red_scratch[wv][tid];
        const int tile_id = (int)blockIdx.x + k * (int)gridDim.x;
        p.partials[(int64_t)tile_id * BACS_NACC + tid] = (double)r;
      }
      consumer_sync();
    }
  }

  // ---- per-CTA partial sums ------------------------------------------------------------------
  if (a.mode != BACS_PIX_SCORE) {
#pragma unroll
    for (int i = 0; i < BACS_NACC; ++i) {
      const float r = warp_sum(acc[i]);
      if (lane == 0) red_scratch[wid][i] = r;
    }
    consumer_sync();
    if (tid < BACS_NACC) {
      double r = 0.0;
      for (int wv = 0; wv < kConsumerWarps; ++wv) r += (double)red_scratch[wv][tid];
      p.partials[(int64_t)blockIdx.x * BACS_NACC + tid] = r;
    }
  }
}

// Deterministic reduction of the partials: block 0 -> acc; in SCORE mode block 1+b -> score[b].
struct PixelEpilogue {
  const int32_t* ready;
  float* focal_scale_out;
  float* loss_out;
  float focal_weight, loss_coef;
  int loss_over_wsum;
};
__global__ void __launch_bounds__(256) pixel_reduce_kernel(const double* __restrict__ partials, int n_part,
                                                           int tiles_per_image, double* __restrict__ acc,
                                                           double* __restrict__ score, double inv_hw,
                                                           const PixelEpilogue ep) {
  __shared__ double scratch[8][BACS_NACC];
  __shared__ double total[BACS_NACC];
  const bool whole = blockIdx.x == 0;
  const int t0 = whole ? 0 : (blockIdx.x - 1) * tiles_per_image;
  const int t1 = whole ? n_part : t0 + tiles_per_image;
  // thread (g, i): accumulator i of partial rows g, g+32, ... (a partial row is 64 contiguous bytes)
  const int i = threadIdx.x & 7, g = threadIdx.x >> 3;
  double s = 0.0;
  pdl_wait();
  pdl_trigger();
  // eight partial rows in flight per thread (one row at a time is a chain of dependent L2 round trips: 6.7 us for the
  // 296 rows of a training step in the ncu launch list); same summation order
  for (int t = t0 + g; t < t1; t += 32 * 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int tt = t + 32 * u;
      v[u] = tt < t1 ? partials[(int64_t)tt * BACS_NACC + i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < 8) scratch[wid][lane] = s;
  __syncthreads();
  if (threadIdx.x < BACS_NACC) {
    double r = 0.0;
    for (int wv = 0; wv < 8; ++wv) r += scratch[wv][threadIdx.x];
    if (whole) {
      acc[threadIdx.x] = r;
      total[threadIdx.x] = r;
    } else if (threadIdx.x == BACS_ACC_LOSS) {
      score[blockIdx.x - 1] = -r * inv_hw;
    }
  }
  if (!whole || (ep.focal_scale_out == nullptr && ep.loss_out == nullptr)) return;
  __syncthreads();
  if (threadIdx.x == 0) {  // the step's scalar epilogue (focal normaliser, loss value)
    const double kept = total[BACS_ACC_KEPT];
    const bool on = (ep.ready == nullptr || *ep.ready != 0) && total[BACS_ACC_BG] > 0.0 && kept > 0.0;
    const double fs = on ? (double)ep.focal_weight / kept : 0.0;
    if (ep.focal_scale_out) *ep.focal_scale_out = (float)fs;
    if (ep.loss_out) {
      double main = (double)ep.loss_coef * total[BACS_ACC_LOSS];
      if (ep.loss_over_wsum) main /= total[BACS_ACC_WSUM];
      *ep.loss_out = (float)(main + fs * total[BACS_ACC_FOCAL]);
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

// [B,K,H,W] as the 3-D tensor (H*W, K, B) with boxes of 256 pixels x K channels x 1 image
static bool make_tile_map(CUtensorMap* map, const void* base, int dtype, int B, int K, int64_t HW, unsigned box_px = 256u) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || K > 256) return false;
  const size_t es = dtype_size(dtype);
  const CUtensorMapDataType dt = dtype == BACS_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == BACS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                      : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)K, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)HW * es, (cuuint64_t)HW * K * es};
  const cuuint32_t box[3] = {box_px, (cuuint32_t)K, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return fn(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// seen logits [B,T,h,w] as the 4-D tensor (w, h, T, B); one box = rows (i0, i0+1) of every head
static bool make_z_map(CUtensorMap* map, const float* z, int B, int T, int h, int w) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || w > 256 || T > 256 || (w * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(z) & 15) != 0) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4, (cuuint64_t)T * h * w * 4};
  const cuuint32_t box[4] = {(cuuint32_t)w, 2u, (cuuint32_t)T, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(z), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_plan(const bacs_pixel_args& a, PixelPlan* plan) {
  const size_t es = dtype_size(a.dtype);
  const int64_t HW = (int64_t)a.H * a.W;
  const int sms = sm_count();
  const size_t extra = (a.z ? (size_t)a.T * a.w * 4 : 0) + (size_t)(a.w + 1) * 4 + 64;
  const size_t cap = 220 * 1024;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };

  // ---- fast path: K <= 24 logits per pixel stay in registers, full 512-pixel tiles, TMA ring ----
  const bool fast_ok = a.K <= 24 && HW % 512 == 0 && (HW * es) % 16 == 0 && aligned16(a.logits) &&
                       aligned16(a.labels) && (!a.dlogits || aligned16(a.dlogits)) && (!a.preds || aligned16(a.preds)) &&
                       (!a.distill_mask || (reinterpret_cast<uintptr_t>(a.distill_mask) & 1) == 0);
  if (fast_ok) {
    static const int kregs[7] = {4, 8, 12, 16, 20, 21, 24};
    int kreg = 24;
    for (int i = 0; i < 7; ++i)
      if (kregs[i] >= a.K) {
        kreg = kregs[i];
        break;
      }
    // 4 stages of {kreg logit rows, 512 int64 labels, seen-head rows} + one [T][8] float2 strip per warp
    const size_t zrows = a.z ? (((size_t)a.T * 2 * a.w * 4 + 127) & ~(size_t)127) : 0;
    const size_t stage = (size_t)kreg * 512 * es + 4096 + zrows;
    const size_t tail = (a.z ? (size_t)8 * a.T * 8 * 8 : 0) + 64;
    // row tiles: a tile lies inside one image row and 64 pixels touch <= 6 low-res columns
    const bool rowtile = a.z != nullptr && a.W % 512 == 0 && a.seen_scale >= 16 && a.w % 4 == 0 &&
                         (reinterpret_cast<uintptr_t>(a.z) & 15) == 0;
    // 228 KB of shared memory per SM, 1 KB reserved per resident CTA, < 0.5 KB static
    auto two_fit = [](size_t smem) { return 2 * (smem + 1024 + 512) <= (size_t)228 * 1024; };
    // the training-step kernel also runs with a 3-stage ring when that is what lets two CTAs share an SM
    const bool wce = rowtile && a.mode == BACS_PIX_WEIGHTED_CE && !a.seen_max && encode_tiled_fn() != nullptr &&
                     a.w <= 256 && a.T <= 256 && a.K <= 256;
    int stages = 4;
    if (wce && !two_fit(4 * stage + tail) && two_fit(3 * stage + tail)) stages = 3;
    const size_t smem = (size_t)stages * stage + tail;
    if (smem + 1024 <= cap) {
      const int per_sm = two_fit(smem) ? 2 : 1;
      const int64_t tiles = HW / 512 * a.B;
      plan->fast = 1;
      plan->coop = 0;
      plan->ppt = 2;
      plan->P = 512;
      plan->kreg = kreg;
      plan->stages = stages;
      plan->smem = smem;
      plan->grid = (int)std::min<int64_t>(tiles, (int64_t)sms * per_sm);
      plan->rowtile = rowtile ? 1 : 0;
      return true;
    }
  }

  // ---- generic path: any K that fits a shared-memory tile, ragged / unaligned shapes ----
  plan->fast = 0;
  plan->kreg = 0;
  for (int min_stages = 2; min_stages >= 1; --min_stages)
    for (int ppt = 2; ppt >= 1; --ppt) {
      if (ppt == 2 && (HW & 1)) continue;  // pixel pairs need even image sizes
      const int P = ppt * kConsumers;
      // one pixel per thread and even images: lane pairs walk the channels together on padded rows (see the kernel)
      const bool coop = ppt == 1 && (HW & 1) == 0 && a.K >= 2;
      const size_t tile = (size_t)a.K * ((size_t)P * es + (coop ? 64 : 0));
      int stages = 0;
      for (int st = 3; st >= min_stages; --st)
        if ((size_t)st * tile + extra + 1024 <= cap) {
          stages = st;
          break;
        }
      if (!stages) continue;
      const int64_t tiles = (HW + P - 1) / P * a.B;
      size_t smem = (size_t)stages * tile + extra;
      int per_sm = 1;
      if (ppt == 1) {  // (measured at K = 151: 2 CTAs x 1 stage 1.57 ms, 1 CTA x 2 stages 1.75 ms)
        // large K: one pixel per thread.  Two CTAs per SM double the warps that hide latency; when two stages each
        // do not fit, one stage each still overlaps -- one CTA computes while the other one loads / stores.
        auto two_fit = [](size_t sm) { return 2 * (sm + 1024 + 512) <= (size_t)228 * 1024; };
        for (int st = std::min(stages, 2); st >= 1; --st)
          if (two_fit((size_t)st * tile + extra)) {
            stages = st;
            smem = (size_t)st * tile + extra;
            per_sm = 2;
            break;
          }
      }
      plan->ppt = ppt;
      plan->coop = coop ? 1 : 0;
      plan->P = P;
      plan->stages = stages;
      plan->smem = smem;
      plan->grid = (int)std::min<int64_t>(tiles, (int64_t)sms * per_sm);
      plan->rowtile = (a.z != nullptr && P <= a.W && a.W % P == 0) ? 1 : 0;
      return true;
    }
  return false;
}

}  // namespace bacs

#include "pixel_lowres.cuh"  // the same loss evaluated from low-res logits (up-sample + adjoint fused)
#include "pixel_stream.cuh"  // large class counts: two streaming passes instead of shared-memory tiles
#include "pixel_regs.cuh"    // large class counts in 16-bit storage: one pass, the channel column in registers

namespace bacs {

struct StreamPlan {
  int blocks_x;
  size_t part_bytes, coef_bytes;
};

// K >= 64 on 16-byte aligned, vector-divisible images (everything else stays on the tile kernel)
static bool make_stream_plan(const bacs_pixel_args& a, StreamPlan* plan) {
  if (getenv("BACS_NO_STREAM")) return false;
  const int64_t HW = (int64_t)a.H * a.W;
  const int n = a.dtype == BACS_F32 ? 4 : 8;
  auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & (m - 1)) == 0; };
  if (a.K < 64 || HW % n != 0 || !al(a.logits, 16) || !al(a.labels, 16)) return false;
  if (a.dlogits && !al(a.dlogits, 16)) return false;
  if (a.preds && !al(a.preds, 16)) return false;
  if (a.distill_mask && !al(a.distill_mask, 4)) return false;
  const int64_t groups = HW / n;
  int64_t bx = (groups + kStreamThreads - 1) / kStreamThreads;
  const int64_t cap = std::max<int64_t>(1, (int64_t)sm_count() * 4);  // per image
  if (bx > cap) bx = cap;
  plan->blocks_x = (int)bx;
  plan->part_bytes = align_up((size_t)bx * a.B * BACS_NACC * sizeof(double), 256);
  plan->coef_bytes = a.dlogits ? align_up((size_t)6 * HW * a.B * sizeof(float), 256) : 0;
  return true;
}

// tile shape of the register-column kernel: 256 threads x 2 CTAs per SM, or (BACS_REGS_CFG=128) 128 threads x 3
static int regs_threads() {
  const char* e = getenv("BACS_REGS_CFG");
  return (e && atoi(e) == 256) ? 256 : 128;
}
static int regs_ctas(int nt) {
  const char* e = getenv("BACS_REGS_CTAS");
  return nt == 256 ? 2 : ((e && atoi(e) == 3) ? 3 : 4);
}

// 16-bit logits with 64 <= K <= kRegsKMax on whole 256-pixel tiles: one pass with the channel column of a pixel pair in
// the registers of two lanes (pixel_regs.cuh)
static bool make_regs_plan(const bacs_pixel_args& a, StreamPlan* plan) {
  if (getenv("BACS_NO_REGS")) return false;
  const int64_t HW = (int64_t)a.H * a.W;
  auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & (m - 1)) == 0; };
  const int nt = regs_threads();
  if (a.dtype == BACS_F32 || a.K < 64 || a.K > kRegsKMax || HW % 256 != 0 || encode_tiled_fn() == nullptr) return false;
  if (a.mode == BACS_PIX_SCORE) return false;  // (per-image sums: streaming path)
  if (!al(a.logits, 16) || !al(a.labels, 8) || (a.dlogits && !al(a.dlogits, 4))) return false;
  const int64_t tiles = HW / nt * a.B;
  if (tiles > 0x3fffffff) return false;
  plan->blocks_x = (int)std::min<int64_t>(tiles, (int64_t)sm_count() * regs_ctas(nt));  // persistent: one wave
  plan->part_bytes = align_up((size_t)plan->blocks_x * BACS_NACC * sizeof(double), 256);
  plan->coef_bytes = 0;
  return true;
}

struct LowresPlan {
  int SX, R, groups, nsrc_max, grid;
  size_t smem, part_bytes, g32_bytes;
};

// host copy of lerp_half_pixel (fp32, op by op)
static void host_half_pixel(int dst, int in_size, float scale, int* i0, int* i1) {
  volatile float t = (float)dst + 0.5f;
  volatile float m = scale * t;
  float src = m - 0.5f;
  if (src < 0.f) src = 0.f;
  *i0 = std::min((int)src, in_size - 1);
  *i1 = *i0 + (*i0 < in_size - 1 ? 1 : 0);
}

static bool make_lowres_plan(const bacs_pixel_args& a, int lh, int lw, LowresPlan* plan) {
  if (lh <= 0 || lw <= 0 || lw > kLowresThreads || a.W % lw != 0 || a.H % lh != 0) return false;
  const int SX = a.W / lw;
  if (SX != 16 && SX != 8) return false;
  if (a.K > 255) return false;
  int R = 1;
  while (2 * R * lw <= kLowresThreads && 2 * R <= a.H) R *= 2;
  const float hy = hp_scale(lh, a.H);
  int groups, nsrc;
  for (;; R /= 2) {  // a row group may touch at most kLowresMaxSrc source rows
    groups = (a.H + R - 1) / R;
    nsrc = 1;
    for (int g = 0; g < groups; ++g) {
      int f0, f1, l0, l1;
      host_half_pixel(g * R, lh, hy, &f0, &f1);
      host_half_pixel(std::min(g * R + R, a.H) - 1, lh, hy, &l0, &l1);
      nsrc = std::max(nsrc, l1 - f0 + 1);
    }
    if (nsrc <= kLowresMaxSrc || R == 1) break;
  }
  plan->SX = SX;
  plan->R = R;
  plan->groups = groups;
  plan->nsrc_max = nsrc;
  plan->grid = groups * a.B;
  plan->smem = ((size_t)nsrc * a.K * (lw + 2) + (size_t)3 * R * kLowresChunk * lw + (size_t)nsrc * R +
                (a.z ? (size_t)R * a.T * a.w : 0) + (size_t)R * (a.w + 1)) * 4 + 16;
  plan->part_bytes = align_up((size_t)plan->grid * BACS_NACC * sizeof(double), 256);
  plan->g32_bytes = (a.dlogits && a.dtype != BACS_F32) ? align_up((size_t)a.B * a.K * lh * lw * 4, 256) : 0;
  return plan->smem <= (size_t)200 * 1024;
}

}  // namespace bacs

using namespace bacs;

extern "C" {

size_t bacs_pixel_lowres_workspace_bytes(const bacs_pixel_args* a, int32_t lh, int32_t lw) {
  if (!a) return 0;
  LowresPlan plan;
  if (!make_lowres_plan(*a, lh, lw, &plan)) return 0;
  return plan.part_bytes + plan.g32_bytes;
}

int bacs_pixel_loss_lowres(const bacs_pixel_args* a, int32_t lh, int32_t lw, void* workspace, size_t workspace_bytes,
                           void* stream) {
  BACS_REQUIRE(a && a->logits && a->labels && a->acc, "bacs_pixel_loss_lowres: null argument");
  BACS_REQUIRE(a->B > 0 && a->K > 0 && a->H > 0 && a->W > 0, "bacs_pixel_loss_lowres: empty shape");
  BACS_REQUIRE(a->dtype == BACS_F32 || a->dtype == BACS_BF16 || a->dtype == BACS_F16, "bacs_pixel_loss_lowres: dtype");
  if (a->z) {
    BACS_REQUIRE(a->T > 0 && a->h > 0 && a->w > 0, "bacs_pixel_loss_lowres: seen logits given without T/h/w");
    BACS_REQUIRE(a->seen_scale > 0 && a->H == a->h * a->seen_scale && a->W == a->w * a->seen_scale,
                 "bacs_pixel_loss_lowres: H,W must equal h,w * seen_scale");
  }
  if (a->gz)
    BACS_REQUIRE(a->z && a->focal_head >= 0 && a->focal_head < a->T, "bacs_pixel_loss_lowres: focal head out of range");
  if (a->mode == BACS_PIX_WEIGHTED_CE)
    BACS_REQUIRE(a->old_cl >= 1 && (a->z || a->seen_max),
                 "bacs_pixel_loss_lowres: WEIGHTED_CE needs old_cl >= 1 and the seen logits / probabilities");
  if (a->mode != BACS_PIX_WEIGHTED_CE && a->dlogits)
    BACS_REQUIRE(a->hist, "bacs_pixel_loss_lowres: CE-type gradients need the label histogram");
  if (a->mode == BACS_PIX_SCORE)
    BACS_REQUIRE(a->score && !a->dlogits, "bacs_pixel_loss_lowres: SCORE mode needs score and no gradient");
  LowresPlan plan;
  if (!make_lowres_plan(*a, lh, lw, &plan)) {
    set_error("bacs_pixel_loss_lowres: unsupported geometry (H=%d W=%d lh=%d lw=%d K=%d): W/lw must be 8 or 16, "
              "lw <= 256, K <= 255", a->H, a->W, lh, lw, a->K);
    return BACS_ERR_UNSUPPORTED;
  }
  if (!workspace || workspace_bytes < plan.part_bytes + plan.g32_bytes) {
    set_error("bacs_pixel_loss_lowres: workspace too small (%zu bytes)", workspace_bytes);
    return BACS_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_low = (int64_t)a->B * a->K * lh * lw;
  LowresParams p;
  p.a = *a;
  p.g32 = nullptr;
  if (a->dlogits) {
    p.g32 = a->dtype == BACS_F32 ? reinterpret_cast<float*>(a->dlogits)
                                 : reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + plan.part_bytes);
    if (cudaMemsetAsync(p.g32, 0, (size_t)n_low * 4, s) != cudaSuccess) {
      set_error("bacs_pixel_loss_lowres: memset failed");
      return BACS_ERR_CUDA;
    }
  }
  p.lh = lh;
  p.lw = lw;
  p.R = plan.R;
  p.groups_per_image = plan.groups;
  p.nsrc_max = plan.nsrc_max;
  p.hy = hp_scale(lh, a->H);
  p.inv_n = (float)(1.0 / ((double)a->B * (double)a->H * (double)a->W));
  p.sy = a->z ? ac_scale(a->h, a->H) : 0.f;
  p.sx = a->z ? ac_scale(a->w, a->W) : 0.f;
  p.partials = reinterpret_cast<double*>(workspace);
#define LAUNCH_LR(SXV)                                                                                          \
  do {                                                                                                          \
    auto kern = pixel_lowres_kernel<SXV>;                                                                       \
    if (plan.smem > 48 * 1024 &&                                                                                \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) != cudaSuccess) { \
      set_error("bacs_pixel_loss_lowres: cannot opt in to %zu bytes of shared memory", plan.smem);              \
      return BACS_ERR_CUDA;                                                                                     \
    }                                                                                                           \
    kern<<<plan.grid, kLowresThreads, plan.smem, s>>>(p);                                                       \
  } while (0)
  if (plan.SX == 16) LAUNCH_LR(16);
  else LAUNCH_LR(8);
#undef LAUNCH_LR
  BACS_CHECK_LAUNCH("bacs_pixel_loss_lowres");
  if (a->dlogits && a->dtype != BACS_F32) {
    const int nb = (int)((n_low + 255) / 256);
    if (a->dtype == BACS_BF16)
      lowres_cast_kernel<__nv_bfloat16><<<nb, 256, 0, s>>>(p.g32, reinterpret_cast<__nv_bfloat16*>(a->dlogits), n_low);
    else
      lowres_cast_kernel<__half><<<nb, 256, 0, s>>>(p.g32, reinterpret_cast<__half*>(a->dlogits), n_low);
    BACS_CHECK_LAUNCH("bacs_pixel_loss_lowres(cast)");
  }
  const int nblk = 1 + (a->mode == BACS_PIX_SCORE ? a->B : 0);
  PixelEpilogue ep;
  ep.ready = a->ready;
  ep.focal_scale_out = a->focal_scale_out;
  ep.loss_out = a->loss_out;
  ep.focal_weight = a->focal_weight;
  ep.loss_coef = a->loss_coef;
  ep.loss_over_wsum = a->loss_over_wsum;
  launch_pdl(pixel_reduce_kernel, dim3(nblk), dim3(256), 0, s, p.partials, plan.grid, plan.groups, a->acc, a->score,
                                           1.0 / ((double)a->H * (double)a->W), ep);
  BACS_CHECK_LAUNCH("bacs_pixel_loss_lowres(reduce)");
  return BACS_OK;
}

size_t bacs_pixel_workspace_bytes(const bacs_pixel_args* a) {
  if (!a) return 0;
  StreamPlan sp;
  if (make_regs_plan(*a, &sp)) return sp.part_bytes;
  if (make_stream_plan(*a, &sp)) return sp.part_bytes + sp.coef_bytes;
  PixelPlan plan;
  if (!make_plan(*a, &plan)) return 0;
  const int64_t HW = (int64_t)a->H * a->W;
  const int64_t tiles = (HW + plan.P - 1) / plan.P * a->B;
  const int64_t n_part = a->mode == BACS_PIX_SCORE ? tiles : plan.grid;
  return align_up((size_t)n_part * BACS_NACC * sizeof(double), 256);
}

static bool wce_eligible(const bacs_pixel_args& a, const PixelPlan& plan) {
  return plan.fast && plan.rowtile && a.mode == BACS_PIX_WEIGHTED_CE && !a.seen_max && a.z && encode_tiled_fn() != nullptr &&
         a.w <= 256 && a.T <= 256 && a.K <= 256;
}

int bacs_pixel_kernel_variant(const bacs_pixel_args* a) {
  PixelPlan plan;
  StreamPlan sp;
  if (a && make_regs_plan(*a, &sp)) return 5;
  if (a && make_stream_plan(*a, &sp)) return 4;
  if (!a || !make_plan(*a, &plan)) return -1;
  return wce_eligible(*a, plan) ? 2 : (plan.fast ? 1 : 0);
}

int bacs_pixel_loss(const bacs_pixel_args* a, void* workspace, size_t workspace_bytes, bacs_stream_t stream) {
  BACS_REQUIRE(a, "bacs_pixel_loss: null args");
  BACS_REQUIRE(a->logits && a->labels && a->acc, "bacs_pixel_loss: logits, labels and acc are required");
  BACS_REQUIRE(a->B > 0 && a->K > 0 && a->H > 0 && a->W > 0, "bacs_pixel_loss: bad shape");
  BACS_REQUIRE(a->ignore_index >= a->K || a->ignore_index < 0, "bacs_pixel_loss: ignore_index inside the class range");
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
  if (a->mode == BACS_PIX_SCORE)
    BACS_REQUIRE(a->score && !a->dlogits, "bacs_pixel_loss: SCORE mode needs score and no gradient");
  StreamPlan sp;
  if (make_regs_plan(*a, &sp)) {  // large K, 16-bit logits: one pass, channel column in registers (pixel_regs.cuh)
    if (!workspace || workspace_bytes < sp.part_bytes) {
      set_error("bacs_pixel_loss: workspace too small (%zu bytes)", workspace_bytes);
      return BACS_ERR_WORKSPACE;
    }
    RegsParams rq;
    StreamParams& q = rq.s;
    q.a = *a;
    q.partials = reinterpret_cast<double*>(workspace);
    q.coef = a->dlogits ? reinterpret_cast<float*>(workspace) : nullptr;  // only "a gradient is wanted" (stream_norm)
    q.blocks_x = sp.blocks_x;
    q.b0 = 0;
    q.inv_n = (float)(1.0 / ((double)a->B * (double)a->H * (double)a->W));
    q.sy = a->z ? ac_scale(a->h, a->H) : 0.f;
    q.sx = a->z ? ac_scale(a->w, a->W) : 0.f;
    const int64_t HWr = (int64_t)a->H * a->W;
    const int nt = regs_threads();
    rq.tiles_per_image = (int)(HWr / nt);
    rq.n_tiles = rq.tiles_per_image * a->B;
    if (!make_tile_map(&rq.tmap_in, a->logits, a->dtype, a->B, a->K, HWr, (unsigned)nt)) {
      set_error("bacs_pixel_loss: cuTensorMapEncodeTiled failed for the logits");
      return BACS_ERR_CUDA;
    }
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)sp.blocks_x);
    const size_t smem = (size_t)a->K * nt * 2 + 128;
#define LAUNCH_REGS(TT)                                                                                         \
  do {                                                                                                          \
    auto kern = nt == 256 ? pixel_regs_kernel<TT, kRegsKH, 256, 2>                                              \
                          : (regs_ctas(nt) == 3 ? pixel_regs_kernel<TT, kRegsKH, 128, 3> : pixel_regs_kernel<TT, kRegsKH, 128, 4>); \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {    \
      set_error("bacs_pixel_loss: cannot opt in to %zu bytes of shared memory", smem);                          \
      return BACS_ERR_CUDA;                                                                                     \
    }                                                                                                           \
    kern<<<grid, nt, smem, st>>>(rq);                                                                 \
  } while (0)
    if (a->dtype == BACS_BF16) LAUNCH_REGS(__nv_bfloat16);
    else LAUNCH_REGS(__half);
#undef LAUNCH_REGS
    BACS_CHECK_LAUNCH("bacs_pixel_loss(register column)");
    const int nblk = 1 + (a->mode == BACS_PIX_SCORE ? a->B : 0);
    PixelEpilogue ep;
    ep.ready = a->ready;
    ep.focal_scale_out = a->focal_scale_out;
    ep.loss_out = a->loss_out;
    ep.focal_weight = a->focal_weight;
    ep.loss_coef = a->loss_coef;
    ep.loss_over_wsum = a->loss_over_wsum;
    launch_pdl(pixel_reduce_kernel, dim3(nblk), dim3(256), 0, st, q.partials, sp.blocks_x, sp.blocks_x, a->acc,
               a->score, 1.0 / ((double)a->H * (double)a->W), ep);
    BACS_CHECK_LAUNCH("bacs_pixel_loss(reduce)");
    return BACS_OK;
  }
  if (make_stream_plan(*a, &sp)) {  // large K: two streaming passes (pixel_stream.cuh)
    if (!workspace || workspace_bytes < sp.part_bytes + sp.coef_bytes) {
      set_error("bacs_pixel_loss: workspace too small (%zu bytes)", workspace_bytes);
      return BACS_ERR_WORKSPACE;
    }
    StreamParams q;
    q.a = *a;
    q.partials = reinterpret_cast<double*>(workspace);
    q.coef = a->dlogits ? reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + sp.part_bytes) : nullptr;
    q.blocks_x = sp.blocks_x;
    q.inv_n = (float)(1.0 / ((double)a->B * (double)a->H * (double)a->W));
    q.sy = a->z ? ac_scale(a->h, a->H) : 0.f;
    q.sx = a->z ? ac_scale(a->w, a->W) : 0.f;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_part = sp.blocks_x * a->B;
    // (launching a few images at a time so that pass B re-reads them from the 126 MB L2 was measured: 1 / 2 / 4 / 8
    //  images per launch pair 2.67 / 1.69 / 1.66 / 1.48 ms against 1.33 ms for all 24 at once -- under-filled launches
    //  cost more than the second HBM read)
    const int per = a->B;
    for (int b0 = 0; b0 < a->B; b0 += per) {
      q.b0 = b0;
      dim3 grid((unsigned)sp.blocks_x, (unsigned)std::min(per, a->B - b0));
      BACS_DISPATCH_DTYPE(a->dtype, TT, { pixel_stream_stats_kernel<TT><<<grid, kStreamThreads, 0, st>>>(q); });
      BACS_CHECK_LAUNCH("bacs_pixel_loss(stream stats)");
      if (a->dlogits) {
        BACS_DISPATCH_DTYPE(a->dtype, TT, { pixel_stream_grad_kernel<TT><<<grid, kStreamThreads, 0, st>>>(q); });
        BACS_CHECK_LAUNCH("bacs_pixel_loss(stream grad)");
      }
    }
    const int nblk = 1 + (a->mode == BACS_PIX_SCORE ? a->B : 0);
    PixelEpilogue ep;
    ep.ready = a->ready;
    ep.focal_scale_out = a->focal_scale_out;
    ep.loss_out = a->loss_out;
    ep.focal_weight = a->focal_weight;
    ep.loss_coef = a->loss_coef;
    ep.loss_over_wsum = a->loss_over_wsum;
    launch_pdl(pixel_reduce_kernel, dim3(nblk), dim3(256), 0, st, q.partials, n_part, sp.blocks_x, a->acc, a->score,
                                              1.0 / ((double)a->H * (double)a->W), ep);
    BACS_CHECK_LAUNCH("bacs_pixel_loss(reduce)");
    return BACS_OK;
  }
  PixelPlan plan;
  if (!make_plan(*a, &plan)) {
    set_error("bacs_pixel_loss: K=%d too large for a shared-memory tile", a->K);
    return BACS_ERR_UNSUPPORTED;
  }
  const int64_t HW = (int64_t)a->H * a->W;
  const int64_t tiles_per_image = (HW + plan.P - 1) / plan.P;
  const int64_t n_tiles = tiles_per_image * a->B;
  BACS_REQUIRE(n_tiles < 0x7fffffff, "bacs_pixel_loss: too many tiles");
  const int64_t n_part = a->mode == BACS_PIX_SCORE ? n_tiles : plan.grid;
  if (!workspace || workspace_bytes < (size_t)n_part * BACS_NACC * sizeof(double)) {
    set_error("bacs_pixel_loss: workspace too small (%zu bytes)", workspace_bytes);
    return BACS_ERR_WORKSPACE;
  }
  const size_t es = dtype_size(a->dtype);
  PixelParams p;
  p.a = *a;
  p.P = plan.P;
  p.tiles_per_image = (int)tiles_per_image;
  p.n_tiles = (int)n_tiles;
  p.stages = plan.stages;
  p.inv_n = (float)(1.0 / ((double)a->B * (double)HW));
  p.sy = a->z ? ac_scale(a->h, a->H) : 0.f;
  p.sx = a->z ? ac_scale(a->w, a->W) : 0.f;
  p.partials = reinterpret_cast<double*>(workspace);
  p.use_tmap = 0;
  if (plan.fast) {
    const bool ok_in = make_tile_map(&p.tmap_in, a->logits, a->dtype, a->B, a->K, HW);
    const bool ok_out = !a->dlogits || make_tile_map(&p.tmap_out, a->dlogits, a->dtype, a->B, a->K, HW);
    const bool ok_z = !(plan.rowtile && a->z) || make_z_map(&p.tmap_z, a->z, a->B, a->T, a->h, a->w);
    p.use_tmap = (ok_in && ok_out && ok_z) ? 1 : 0;
  }
  p.use_bulk = (((HW * es) % 16 == 0) && ((reinterpret_cast<uintptr_t>(a->logits) & 15) == 0) &&
                (!a->dlogits || (reinterpret_cast<uintptr_t>(a->dlogits) & 15) == 0))
                   ? 1
                   : 0;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_PIX(TT, PPT, KREG, ROWT, COOPV)                                                                 \
  do {                                                                                                         \
    auto kern = pixel_loss_kernel<TT, PPT, KREG, ROWT, COOPV>;                                                        \
    if (plan.smem > 48 * 1024) {                                                                               \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem); \
      if (e != cudaSuccess) {                                                                                  \
        set_error("bacs_pixel_loss: cannot opt in to %zu bytes of shared memory: %s", plan.smem,              \
                  cudaGetErrorString(e));                                                                      \
        return BACS_ERR_CUDA;                                                                                  \
      }                                                                                                        \
    }                                                                                                          \
    kern<<<plan.grid, kConsumers + 32, plan.smem, s>>>(p);                                                     \
  } while (0)
#define LAUNCH_PIX_K(TT, PPT, COOPV)                       \
  do {                                                     \
    if (plan.rowtile) LAUNCH_PIX(TT, PPT, 0, true, COOPV); \
    else LAUNCH_PIX(TT, PPT, 0, false, COOPV);             \
  } while (0)
  const bool wce = p.use_tmap && wce_eligible(*a, plan);
  if (plan.fast && plan.stages != 4 && !wce) {
    set_error("bacs_pixel_loss: tensor-map creation failed for the 3-stage training kernel");
    return BACS_ERR_CUDA;
  }
  if (wce) {  // the training step's kernel
    int rc;
    switch (a->dtype) {
      case BACS_F32: rc = launch_pixel_wce_f32(p, plan, s); break;
      case BACS_BF16: rc = launch_pixel_wce_bf16(p, plan, s); break;
      default: rc = launch_pixel_wce_f16(p, plan, s); break;
    }
    if (rc != BACS_OK) return rc;
  } else if (plan.fast) {
    int rc;
    switch (a->dtype) {
      case BACS_F32: rc = launch_pixel_fast_f32(p, plan, s); break;
      case BACS_BF16: rc = launch_pixel_fast_bf16(p, plan, s); break;
      default: rc = launch_pixel_fast_f16(p, plan, s); break;
    }
    if (rc != BACS_OK) return rc;
  } else {
    BACS_DISPATCH_DTYPE(a->dtype, TT, {
      if (plan.ppt == 2) LAUNCH_PIX_K(TT, 2, false);
      else if (plan.coop) LAUNCH_PIX_K(TT, 1, true);
      else LAUNCH_PIX_K(TT, 1, false);
    });
  }
#undef LAUNCH_PIX_K
#undef LAUNCH_PIX
  BACS_CHECK_LAUNCH("bacs_pixel_loss");
  const int nblk = 1 + (a->mode == BACS_PIX_SCORE ? a->B : 0);
  PixelEpilogue ep;
  ep.ready = a->ready;
  ep.focal_scale_out = a->focal_scale_out;
  ep.loss_out = a->loss_out;
  ep.focal_weight = a->focal_weight;
  ep.loss_coef = a->loss_coef;
  ep.loss_over_wsum = a->loss_over_wsum;
  launch_pdl(pixel_reduce_kernel, dim3(nblk), dim3(256), 0, s, p.partials, (int)n_part, p.tiles_per_image, a->acc, a->score,
                                           1.0 / (double)HW, ep);
  BACS_CHECK_LAUNCH("bacs_pixel_loss(reduce)");
  return BACS_OK;
}

}  // extern "C"
