// Teacher distillation on the last attention map (loss/bacs_loss.py:258-294), fused
// forward + backward without ever materialising the [B,A,H,W] up-sampled tensors.
//
//   L = coef * sum_{b,a,Y} sqrt( S[b,a,Y] ),   S = sum_X ( m * (U(old)^2 - U(new)^2) )^2
//   U = bilinear up-sample (align_corners=False) of the [h,w] attention map.
//
// Inside one low-res cell j of row Y the up-sampled value is linear in the x-weight:
//   U = a + tau * d  (a = value at the cell centre, d = r[j+1] - r[j], tau = lambda - 1/2)
// so U(old)^2 - U(new)^2 = c0 + c1 tau + c2 tau^2 and the masked sum over the cell's pixels
// of its square needs only the five mask moments  M_k = sum_X m * tau^k  (k = 0..4), which
// do not depend on the channel:
//   S_cell = c0^2 M0 + 2 c0 c1 M1 + (c1^2 + 2 c0 c2) M2 + 2 c1 c2 M3 + c2^2 M4 .
// This is algebraically identical to the reference and cuts the work per (channel,row) from
// W pixel evaluations to w cell evaluations (16x fewer for the stride-16 networks).  The
// centring of tau keeps the fp32 error at the level of the direct evaluation.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace bacs {

struct DistillTables {
  int* yi0;       // [H] source row of every full-res row
  float* ywt;     // [H] weight of row yi0+1
  int* xi0;       // [W]
  float* xwt;     // [W]
  int* rowstart;  // [h+1] first full-res row of every source-row interval
  int* colstart;  // [w+1]
};

__global__ void __launch_bounds__(256) distill_tables_kernel(DistillTables t, int H, int W, int h, int w, float sy,
                                                             float sx) {
  for (int Y = threadIdx.x; Y < H; Y += blockDim.x) {
    const Lerp l = lerp_half_pixel(Y, h, sy);
    t.yi0[Y] = l.i0;
    t.ywt[Y] = (l.i1 == l.i0) ? 0.f : l.w1;
  }
  for (int X = threadIdx.x; X < W; X += blockDim.x) {
    const Lerp l = lerp_half_pixel(X, w, sx);
    t.xi0[X] = l.i0;
    t.xwt[X] = (l.i1 == l.i0) ? 0.f : l.w1;
  }
  __syncthreads();
  for (int Y = threadIdx.x; Y < H; Y += blockDim.x)
    if (Y == 0 || t.yi0[Y] != t.yi0[Y - 1]) t.rowstart[t.yi0[Y]] = Y;
  for (int X = threadIdx.x; X < W; X += blockDim.x)
    if (X == 0 || t.xi0[X] != t.xi0[X - 1]) t.colstart[t.xi0[X]] = X;
  if (threadIdx.x == 0) {
    t.rowstart[h] = H;
    t.colstart[w] = W;
  }
}

// moments[((b*H + Y)*5 + k)*w + j] = sum over the pixels X of cell j of m * (lambda - 1/2)^k
__global__ void __launch_bounds__(256) distill_moments_kernel(const uint8_t* __restrict__ mask, int B, int H, int W,
                                                              int w, const int* __restrict__ colstart,
                                                              const float* __restrict__ xwt,
                                                              float* __restrict__ moments) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * w) return;
  const int j = (int)(idx % w);
  const int64_t row = idx / w;  // b*H + Y
  const int x0 = colstart[j], x1 = colstart[j + 1];
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
  const uint8_t* mrow = mask ? mask + row * W : nullptr;
  auto add = [&](int X) {
    const float t = xwt[X] - 0.5f;
    const float t2 = t * t;
    m0 += 1.f;
    m1 += t;
    m2 += t2;
    m3 += t2 * t;
    m4 += t2 * t2;
  };
  if (mrow && ((reinterpret_cast<uintptr_t>(mrow + x0) & 7) == 0) && ((x1 - x0) & 7) == 0) {
    for (int X = x0; X < x1; X += 8) {  // 8 mask bytes per load
      const uint2 v = *reinterpret_cast<const uint2*>(mrow + X);
      if ((v.x | v.y) == 0) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((v.x >> (8 * k)) & 0xffu) add(X + k);
        if ((v.y >> (8 * k)) & 0xffu) add(X + 4 + k);
      }
    }
  } else {
    for (int X = x0; X < x1; ++X)
      if (mrow == nullptr || mrow[X]) add(X);
  }
  float* out = moments + row * 5 * w + j;
  out[0] = m0;
  out[w] = m1;
  out[2 * w] = m2;
  out[3 * w] = m3;
  out[4 * w] = m4;
}

#ifndef BACS_DISTILL_RB
#define BACS_DISTILL_RB 4
#endif
constexpr int kDistillWarps = 8;
// channels per warp: 2 when a lane owns one low-res column (w <= 32), 1 for wider maps (register budget)
__host__ __device__ constexpr int distill_nc(int cpl) { return cpl == 1 ? 2 : 1; }

// Value of a quantity that is linear in the row weight ty:  v(ty) = p + q * ty.
struct Lin {
  float p, q;
  __device__ __forceinline__ float at(float ty) const { return fmaf(q, ty, p); }
};
__device__ __forceinline__ Lin lin(float v0, float v1) { return Lin{v0, v1 - v0}; }  // v0 at ty=0, v1 at ty=1

__device__ __forceinline__ void dist_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void dist_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void dist_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t b32 = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"(b32)
               : "memory");
}
__device__ __forceinline__ void dist_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DWAIT_DONE;\n"
      "bra DWAIT_LOOP;\n"
      "DWAIT_DONE:\n"
      "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ float rsqrt_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One warp owns kChanPerWarp channels of one image; lane j owns low-res column(s) j (+32m).
// The CTA walks the source-row intervals top to bottom; the five mask moments of every
// row of the interval arrive by TMA bulk copy one interval ahead and are shared by all channels.
//
// Per (channel, row, cell), with a = value at the cell centre and d = r[j+1]-r[j] of the
// y-interpolated row (all linear in the row weight ty):
//   alpha = a_o - a_n, sigma = a_o + a_n, delta = d_o - d_n, eps = d_o + d_n
//   U(old)^2 - U(new)^2 = (alpha + delta tau)(sigma + eps tau) = c0 + c1 tau + c2 tau^2,
//   c0 = alpha sigma, c1 = alpha eps + delta sigma, c2 = delta eps            (exactly 0 for old == new)
//   u = G c (G = Hankel matrix of the moments), S_cell = c.u, dS/dc = 2u
//   dL/da_n = -rs (u0 2a_n + u1 2d_n), dL/dd_n = -rs (u1 2a_n + u2 2d_n), 2a_n = sigma - alpha, 2d_n = eps - delta,
//   accumulated against {1, ty} so that the corner gradients are assembled once per interval.
template <typename T, int CPL>
__global__ void __launch_bounds__(32 * kDistillWarps, (CPL <= 2 ? 2 : 1)) distill_kernel(const T* __restrict__ old_att,
                                                                      const T* __restrict__ new_att, int A, int h,
                                                                      int w, int H, DistillTables tb,
                                                                      const float* __restrict__ moments, int rows_max,
                                                                      float grad_coef, T* __restrict__ dnew,
                                                                      double* __restrict__ partials, int n_groups,
                                                                      int total_units) {
  // Work unit = (image b, group of kChanPerCta channels, source-row interval i), flattened in that order.
  // Every persistent CTA takes one contiguous, equally long range of units, so the SMs stay evenly loaded
  // whatever B*A is.  A range that starts inside an image first replays the interval above it ("halo") to
  // obtain the gradient that interval sends down to the range's first source row.
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool want_grad = dnew != nullptr;
  const int u0 = (int)((int64_t)blockIdx.x * total_units / gridDim.x);
  const int u1 = (int)((int64_t)(blockIdx.x + 1) * total_units / gridDim.x);
  const int ustart = u0 - ((want_grad && (u0 % h) != 0) ? 1 : 0);
  // dynamic smem: 3 x [rows_max][5][WP] moment buffers (columns zero-padded to WP) | ty[H] | rowstart[h+1]
  extern __shared__ __align__(16) float s_dyn[];
  __shared__ double red_scratch[32];
  __shared__ uint64_t mom_bar[3];   // moments of an interval have landed (TMA expect-tx)
  __shared__ uint64_t free_bar[3];  // all warps are done with the buffer
  constexpr int WP = 32 * CPL;
  constexpr int NC = distill_nc(CPL);
  constexpr int kChanPerWarp = NC, kChanPerCta = kDistillWarps * NC;
  constexpr unsigned kFull = 0xffffffffu;
  const size_t mom_stride = (size_t)rows_max * 5 * WP;
  float* s_ty = s_dyn + 3 * mom_stride;
  int* s_rowstart = reinterpret_cast<int*>(s_ty + H);
  const bool bulk_ok = (w == WP);  // contiguous rows -> one TMA bulk copy per interval
  for (int k = threadIdx.x; k < H; k += blockDim.x) s_ty[k] = tb.ywt[k];
  for (int k = threadIdx.x; k <= h; k += blockDim.x) s_rowstart[k] = tb.rowstart[k];
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) {
      dist_mbar_init(&mom_bar[k], 1);
      dist_mbar_init(&free_bar[k], kDistillWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto stage_moments = [&](int un, int buf) {  // moments of the rows of unit un -> buffer buf
    const int g = un / h, i = un - g * h, b = g / n_groups;
    const int Y0 = s_rowstart[i], nr = s_rowstart[i + 1] - Y0;
    const float* src = moments + ((int64_t)b * H + Y0) * 5 * w;
    float* dst = s_dyn + buf * mom_stride;
    if (bulk_ok) {
      if (threadIdx.x == 0) dist_bulk_load(dst, src, (uint32_t)(nr * 5 * w * sizeof(float)), &mom_bar[buf]);
    } else {
      const int n = nr * 5 * WP;
      for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int rk = k / WP, j = k - rk * WP;
        dst[k] = j < w ? __ldg(src + rk * w + j) : 0.f;
      }
    }
  };

  // attention values of this lane's column(s): upper source row of the current interval (prev) and, still in
  // storage precision, the row after it (nxt): it is fetched one interval ahead and only converted when it
  // becomes the current row, so the loads have a whole interval to land.  The right neighbour comes by shuffle.
  float o_prev[NC][CPL], n_prev[NC][CPL];
  T o_nxt[NC][CPL], n_nxt[NC][CPL];
  float carry[NC][CPL];  // gradient already collected for the upper row by the interval above
  float loss_acc = 0.f;   // row norms known to every lane (scalar tail rows): counted once per warp
  float loss_quad = 0.f;  // row norms held by the 4 lanes of a quad (transposed reduction)
  auto load_row = [&](const T* base, int b, int ch, int row, T* v) {
#pragma unroll
    for (int m = 0; m < CPL; ++m) {
      const int j = lane + 32 * m;
      v[m] = (j < w && ch < A) ? base[(((int64_t)b * A + ch) * h + row) * w + j] : DT<T>::from_f(0.f);
    }
  };
  // value of column j+1 (the last column is its own right neighbour)
  auto right = [&](const float* v, float* v1) {
#pragma unroll
    for (int m = 0; m < CPL; ++m) {
      const float dn = __shfl_down_sync(kFull, v[m], 1);
      const float wrap = __shfl_sync(kFull, m + 1 < CPL ? v[m + 1 < CPL ? m + 1 : m] : 0.f, 0);
      const int j = lane + 32 * m;
      v1[m] = (j >= w - 1) ? v[m] : (lane == 31 ? wrap : dn);
    }
  };
  if (u0 < u1) stage_moments(ustart, 0);
  if (!bulk_ok) __syncthreads();

  for (int un = ustart, k = 0; un < u1; ++un, ++k) {
    const int g = un / h, i = un - g * h;
    const int b = g / n_groups;
    const int ch0 = (g - b * n_groups) * kChanPerCta + wid * kChanPerWarp;
    const bool halo = un < u0;  // replayed for the gradient it sends down only: no loss, no dnew row
    const int Y0 = s_rowstart[i], Y1 = s_rowstart[i + 1];
    const int nrows = Y1 - Y0;
    const int row1 = min(i + 1, h - 1);
    const int buf = k % 3;
    const float* s_mom = s_dyn + buf * mom_stride;
    const float loss_before = loss_acc, quad_before = loss_quad;
    // Unit k+1 goes into the buffer unit k-2 used: its readers had a whole unit to finish, so the staging
    // thread practically never waits and no CTA-wide barrier is needed.
    if (bulk_ok) {
      if (threadIdx.x == 0 && un + 1 < u1) {
        if (k >= 2) dist_mbar_wait(&free_bar[(k + 1) % 3], (uint32_t)(((k - 2) / 3) & 1));
        stage_moments(un + 1, (k + 1) % 3);
      }
      __syncwarp();
      dist_mbar_wait(&mom_bar[buf], (uint32_t)((k / 3) & 1));
    } else {
      if (un + 1 < u1) stage_moments(un + 1, (k + 1) % 3);  // for the NEXT unit; published by the barrier below
    }
    if (k == 0 || i == 0) {  // first unit of a (image, channel group) segment: fetch its two source rows
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        load_row(old_att, b, ch0 + c, i, o_nxt[c]);
        load_row(new_att, b, ch0 + c, i, n_nxt[c]);
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          o_prev[c][m] = DT<T>::to_f(o_nxt[c][m]);
          n_prev[c][m] = DT<T>::to_f(n_nxt[c][m]);
          carry[c][m] = 0.f;
        }
        load_row(old_att, b, ch0 + c, row1, o_nxt[c]);
        load_row(new_att, b, ch0 + c, row1, n_nxt[c]);
      }
    }

    Lin alpha[NC][CPL], sigma[NC][CPL], delta[NC][CPL], eps[NC][CPL];
    float GA0[NC][CPL], GA1[NC][CPL], GD0[NC][CPL], GD1[NC][CPL];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float o_cur[CPL], n_cur[CPL], o_p1[CPL], n_p1[CPL], o_c1[CPL], n_c1[CPL];
#pragma unroll
      for (int m = 0; m < CPL; ++m) {
        o_cur[m] = row1 != i ? DT<T>::to_f(o_nxt[c][m]) : o_prev[c][m];
        n_cur[m] = row1 != i ? DT<T>::to_f(n_nxt[c][m]) : n_prev[c][m];
      }
      if (i + 2 < h) {  // prefetch the row of the interval after the next one
        load_row(old_att, b, ch0 + c, i + 2, o_nxt[c]);
        load_row(new_att, b, ch0 + c, i + 2, n_nxt[c]);
      }
      right(o_prev[c], o_p1);
      right(n_prev[c], n_p1);
      right(o_cur, o_c1);
      right(n_cur, n_c1);
#pragma unroll
      for (int m = 0; m < CPL; ++m) {
        // values at ty = 0 (upper source row) and ty = 1 (lower source row)
        const float ao0 = 0.5f * (o_prev[c][m] + o_p1[m]), ao1 = 0.5f * (o_cur[m] + o_c1[m]);
        const float an0 = 0.5f * (n_prev[c][m] + n_p1[m]), an1 = 0.5f * (n_cur[m] + n_c1[m]);
        const float do0 = o_p1[m] - o_prev[c][m], do1 = o_c1[m] - o_cur[m];
        const float dn0 = n_p1[m] - n_prev[c][m], dn1 = n_c1[m] - n_cur[m];
        alpha[c][m] = lin(ao0 - an0, ao1 - an1);
        sigma[c][m] = lin(ao0 + an0, ao1 + an1);
        delta[c][m] = lin(do0 - dn0, do1 - dn1);
        eps[c][m] = lin(do0 + dn0, do1 + dn1);
        GA0[c][m] = GA1[c][m] = GD0[c][m] = GD1[c][m] = 0.f;
        o_prev[c][m] = o_cur[m];
        n_prev[c][m] = n_cur[m];
      }
    }

    // Rows are processed four at a time as two PAIRS of adjacent rows: every fp32 operation of a pair is one
    // packed instruction (FFMA2 / FMUL2 / FADD2), with the channel's coefficients as broadcast operands and
    // the rows' moments / weights in the two halves.  A 1-row scalar tail handles the remaining rows.
    F2 GA0p[NC][CPL], GA1p[NC][CPL], GD0p[NC][CPL], GD1p[NC][CPL];  // halves: even / odd row of the pair
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int m = 0; m < CPL; ++m) GA0p[c][m] = GA1p[c][m] = GD0p[c][m] = GD1p[c][m] = f2b(0.f);
    auto process_pairs = [&](const float* mrow, const float* tyrow) {
      constexpr int RP = 2;
      F2 ty2[RP];
      F2 u0[RP][NC][CPL], u1[RP][NC][CPL], u2[RP][NC][CPL], va2[RP][NC][CPL], vd2[RP][NC][CPL], S[RP][NC];
#pragma unroll
      for (int q = 0; q < RP; ++q) {
        ty2[q] = f2(tyrow[2 * q], tyrow[2 * q + 1]);
        const float* mra = mrow + (2 * q) * 5 * WP;
        const float* mrb = mra + 5 * WP;
        F2 M0[CPL], M1[CPL], M2[CPL], M3[CPL], M4[CPL];
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          M0[m] = f2(mra[32 * m], mrb[32 * m]);
          M1[m] = f2(mra[WP + 32 * m], mrb[WP + 32 * m]);
          M2[m] = f2(mra[2 * WP + 32 * m], mrb[2 * WP + 32 * m]);
          M3[m] = f2(mra[3 * WP + 32 * m], mrb[3 * WP + 32 * m]);
          M4[m] = f2(mra[4 * WP + 32 * m], mrb[4 * WP + 32 * m]);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#pragma unroll
          for (int m = 0; m < CPL; ++m) {
            const F2 al = fma2(f2b(alpha[c][m].q), ty2[q], f2b(alpha[c][m].p));
            const F2 sg = fma2(f2b(sigma[c][m].q), ty2[q], f2b(sigma[c][m].p));
            const F2 de = fma2(f2b(delta[c][m].q), ty2[q], f2b(delta[c][m].p));
            const F2 ep = fma2(f2b(eps[c][m].q), ty2[q], f2b(eps[c][m].p));
            va2[q][c][m] = sub2(sg, al);  // 2 a_n
            vd2[q][c][m] = sub2(ep, de);  // 2 d_n
            const F2 c0 = mul2(al, sg);
            const F2 c1 = fma2(al, ep, mul2(de, sg));
            const F2 c2 = mul2(de, ep);
            u0[q][c][m] = fma2(c2, M2[m], fma2(c1, M1[m], mul2(c0, M0[m])));
            u1[q][c][m] = fma2(c2, M3[m], fma2(c1, M2[m], mul2(c0, M1[m])));
            u2[q][c][m] = fma2(c2, M4[m], fma2(c1, M3[m], mul2(c0, M2[m])));
            const F2 sc = fma2(c2, u2[q][c][m], fma2(c1, u1[q][c][m], mul2(c0, u0[q][c][m])));
            S[q][c] = m == 0 ? sc : add2(S[q][c], sc);
          }
        }
      }
      // Transposed butterfly over the 8 row sums of this call (2 row pairs x 2 channels x 2 halves): every level
      // halves the number of sums a lane still carries, so 7 + 2 shuffles replace 5 x 8, and ONE rsqrt per lane
      // serves all eight rows.  Afterwards lane l holds the total of row sum (l >> 2) & 7.
      float rs_mine = 0.f;
      float rs_all[RP][NC][2];
      if constexpr (NC == 2) {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
        F2 k0, k1;
        {
          const F2 send0 = up16 ? S[0][0] : S[1][0], send1 = up16 ? S[0][1] : S[1][1];
          const F2 keep0 = up16 ? S[1][0] : S[0][0], keep1 = up16 ? S[1][1] : S[0][1];
          F2 r0, r1;
          r0.v = __shfl_xor_sync(kFull, send0.v, 16);
          r1.v = __shfl_xor_sync(kFull, send1.v, 16);
          k0 = add2(keep0, r0);
          k1 = add2(keep1, r1);
        }
        F2 kk;
        {
          const F2 send = up8 ? k0 : k1, keep = up8 ? k1 : k0;
          F2 r;
          r.v = __shfl_xor_sync(kFull, send.v, 8);
          kk = add2(keep, r);
        }
        float tot;
        {
          const float lo = f2lo(kk), hi = f2hi(kk);
          const float send = up4 ? lo : hi, keep = up4 ? hi : lo;
          tot = keep + __shfl_xor_sync(kFull, send, 4);
        }
        tot += __shfl_xor_sync(kFull, tot, 2);
        tot += __shfl_xor_sync(kFull, tot, 1);
        // zero (or rounding-negative) row: norm 0 and sub-gradient 0, as torch's norm backward
        rs_mine = tot > 0.f ? rsqrt_fast(tot) : 0.f;
        loss_quad = fmaf(tot, rs_mine, loss_quad);  // every row sum is held by 4 lanes: scaled by 1/4 at the end
      } else {  // one channel per warp: plain butterfly, every lane ends up with every row sum
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int q = 0; q < RP; ++q)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              F2 other;
              other.v = __shfl_xor_sync(kFull, S[q][c].v, o);
              S[q][c] = add2(S[q][c], other);
            }
#pragma unroll
        for (int q = 0; q < RP; ++q)
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float sa = f2lo(S[q][c]), sb = f2hi(S[q][c]);
            rs_all[q][c][0] = sa > 0.f ? rsqrt_fast(sa) : 0.f;
            rs_all[q][c][1] = sb > 0.f ? rsqrt_fast(sb) : 0.f;
            loss_acc = fmaf(sa, rs_all[q][c][0], loss_acc);
            loss_acc = fmaf(sb, rs_all[q][c][1], loss_acc);
          }
      }
      if (want_grad) {
#pragma unroll
        for (int q = 0; q < RP; ++q) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            float rsa, rsb;
            if constexpr (NC == 2) {
              rsa = __shfl_sync(kFull, rs_mine, q * 16 + c * 8);
              rsb = __shfl_sync(kFull, rs_mine, q * 16 + c * 8 + 4);
            } else {
              rsa = rs_all[q][c][0];
              rsb = rs_all[q][c][1];
            }
            const F2 rs = f2(rsa, rsb);
            const F2 rst = mul2(rs, ty2[q]);
#pragma unroll
            for (int m = 0; m < CPL; ++m) {
              const F2 xA = fma2(u1[q][c][m], vd2[q][c][m], mul2(u0[q][c][m], va2[q][c][m]));
              const F2 xD = fma2(u2[q][c][m], vd2[q][c][m], mul2(u1[q][c][m], va2[q][c][m]));
              GA0p[c][m] = fma2(rs, xA, GA0p[c][m]);
              GA1p[c][m] = fma2(rst, xA, GA1p[c][m]);
              GD0p[c][m] = fma2(rs, xD, GD0p[c][m]);
              GD1p[c][m] = fma2(rst, xD, GD1p[c][m]);
            }
          }
        }
      }
    };
    auto process_row = [&](const float* mr, float ty) {
      float M0[CPL], M1[CPL], M2[CPL], M3[CPL], M4[CPL];
      float u0[NC][CPL], u1[NC][CPL], u2[NC][CPL], va2[NC][CPL], vd2[NC][CPL], S[NC];
#pragma unroll
      for (int m = 0; m < CPL; ++m) {
        M0[m] = mr[32 * m];
        M1[m] = mr[WP + 32 * m];
        M2[m] = mr[2 * WP + 32 * m];
        M3[m] = mr[3 * WP + 32 * m];
        M4[m] = mr[4 * WP + 32 * m];
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const float al = alpha[c][m].at(ty), sg = sigma[c][m].at(ty);
          const float de = delta[c][m].at(ty), ep = eps[c][m].at(ty);
          va2[c][m] = sg - al;
          vd2[c][m] = ep - de;
          const float c0 = al * sg;
          const float c1 = fmaf(al, ep, de * sg);
          const float c2 = de * ep;
          u0[c][m] = fmaf(c2, M2[m], fmaf(c1, M1[m], c0 * M0[m]));
          u1[c][m] = fmaf(c2, M3[m], fmaf(c1, M2[m], c0 * M1[m]));
          u2[c][m] = fmaf(c2, M4[m], fmaf(c1, M3[m], c0 * M2[m]));
          const float sc = fmaf(c2, u2[c][m], fmaf(c1, u1[c][m], c0 * u0[c][m]));
          S[c] = m == 0 ? sc : S[c] + sc;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < NC; ++c) S[c] += __shfl_xor_sync(kFull, S[c], o);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float rs = S[c] > 0.f ? rsqrt_fast(S[c]) : 0.f;
        loss_acc = fmaf(S[c], rs, loss_acc);
        if (want_grad) {
#pragma unroll
          for (int m = 0; m < CPL; ++m) {
            const float gA = rs * fmaf(u1[c][m], vd2[c][m], u0[c][m] * va2[c][m]);
            const float gD = rs * fmaf(u2[c][m], vd2[c][m], u1[c][m] * va2[c][m]);
            GA0[c][m] += gA;
            GA1[c][m] = fmaf(gA, ty, GA1[c][m]);
            GD0[c][m] += gD;
            GD1[c][m] = fmaf(gD, ty, GD1[c][m]);
          }
        }
      }
    };
    {
      const float* mrow = s_mom + lane;  // moments of the current row, this lane's column
      const float* tyrow = s_ty + Y0;
      int r0 = 0;
      for (; r0 + 4 <= nrows; r0 += 4, mrow += 4 * 5 * WP, tyrow += 4) process_pairs(mrow, tyrow);
      for (; r0 < nrows; ++r0, mrow += 5 * WP, ++tyrow) process_row(mrow, *tyrow);
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          GA0[c][m] += f2lo(GA0p[c][m]) + f2hi(GA0p[c][m]);
          GA1[c][m] += f2lo(GA1p[c][m]) + f2hi(GA1p[c][m]);
          GD0[c][m] += f2lo(GD0p[c][m]) + f2hi(GD0p[c][m]);
          GD1[c][m] += f2lo(GD1p[c][m]) + f2hi(GD1p[c][m]);
        }
    }

    if (bulk_ok) {
      __syncwarp();
      if (lane == 0) dist_mbar_arrive(&free_bar[buf]);  // this warp is done with the interval's moments
    } else {
      __syncthreads();
    }

    if (want_grad) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int ch = ch0 + c;
        // a_n(ty) = an0 + (an1-an0) ty, d_n likewise; an0 = (n0[j]+n0[j+1])/2, dn0 = n0[j+1]-n0[j]
        // (index 0 = upper source row, 1 = lower).  GA / GD were accumulated with 2a_n, 2d_n: factor -1.
        float g0[CPL], g1[CPL], g0n[CPL], g1n[CPL];
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const float dA0 = -(GA0[c][m] - GA1[c][m]), dA1 = -GA1[c][m];  // d/d an(ty=0), d/d an(ty=1)
          const float dD0 = -(GD0[c][m] - GD1[c][m]), dD1 = -GD1[c][m];
          g0[m] = 0.5f * dA0 - dD0;   // upper row, column j
          g0n[m] = 0.5f * dA0 + dD0;  // upper row, column j+1
          g1[m] = 0.5f * dA1 - dD1;   // lower row, column j
          g1n[m] = 0.5f * dA1 + dD1;  // lower row, column j+1
        }
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const int j = lane + 32 * m;
          // lane 0 of group m takes the value of lane 31 of group m-1 (all lanes run both shuffles)
          const float wrap0 = __shfl_sync(kFull, m > 0 ? g0n[m > 0 ? m - 1 : 0] : 0.f, 31);
          const float wrap1 = __shfl_sync(kFull, m > 0 ? g1n[m > 0 ? m - 1 : 0] : 0.f, 31);
          const float sh0 = __shfl_up_sync(kFull, g0n[m], 1);
          const float sh1 = __shfl_up_sync(kFull, g1n[m], 1);
          const float up0 = lane == 0 ? wrap0 : sh0;
          const float up1 = lane == 0 ? wrap1 : sh1;
          if (j > 0 && j < w) {
            g0[m] += up0;
            g1[m] += up1;
          }
          if (j == w - 1) {  // the last column is its own right neighbour
            g0[m] += g0n[m];
            g1[m] += g1n[m];
          }
        }
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const int j = lane + 32 * m;
          float fin = carry[c][m] + g0[m];
          if (row1 == i) fin += g1[m];
          if (!halo && j < w && ch < A) dnew[(((int64_t)b * A + ch) * h + i) * w + j] = DT<T>::from_f(grad_coef * fin);
          carry[c][m] = g1[m];
        }
      }
    }
    if (halo) {
      loss_acc = loss_before;
      loss_quad = quad_before;
    }
  }
  // every lane holds the same loss_acc (the shuffle reduction broadcasts); count it once per warp
  const double mine = ((lane == 0) ? (double)loss_acc : 0.0) + 0.25 * (double)loss_quad;
  const double tot = block_sum(mine, red_scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(const double* __restrict__ partials, int n,
                                                            double* __restrict__ out, float coef,
                                                            float* __restrict__ out_scaled,
                                                            const float* __restrict__ addend) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    out[0] = s;
    if (out_scaled) out_scaled[0] = (float)((double)coef * s) + (addend ? addend[0] : 0.f);
  }
}

struct DistillLayout {
  size_t off_yi0, off_ywt, off_xi0, off_xwt, off_rowstart, off_colstart, off_moments, off_partials, total;
  int n_cta_x;
};
static DistillLayout distill_layout(int B, int A, int h, int w, int H, int W) {
  DistillLayout l;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  l.off_yi0 = take(sizeof(int) * H);
  l.off_ywt = take(sizeof(float) * H);
  l.off_xi0 = take(sizeof(int) * W);
  l.off_xwt = take(sizeof(float) * W);
  l.off_rowstart = take(sizeof(int) * (h + 1));
  l.off_colstart = take(sizeof(int) * (w + 1));
  l.off_moments = take(sizeof(float) * (size_t)B * H * 5 * w);
  const int chan_per_cta = kDistillWarps * distill_nc(w <= 32 ? 1 : (w <= 64 ? 2 : 4));
  l.n_cta_x = (A + chan_per_cta - 1) / chan_per_cta;
  l.off_partials = take(sizeof(double) * (size_t)std::max(2 * sm_count(), 1024));
  l.total = o;
  return l;
}

// tensor-core path (distill_tc.cu)
bool distill_tc_shape_ok(int dtype, int B, int A, int h, int w, int H, int W);
size_t distill_tc_workspace_bytes(int dtype, int B, int A, int h, int w, int H, int W);
int distill_tc_launch(const void* old_att, const void* new_att, int dtype, int B, int A, int h, int w, const uint8_t* mask, int H,
                      int W, float grad_coef, double* loss_sum, float* loss_scaled, const float* addend, void* dnew,
                      void* workspace, size_t workspace_bytes, cudaStream_t s);
static int g_distill_mode = 0;  // 0: tensor cores where they apply, 1: FMA kernel only, 2: tensor cores or error

}  // namespace bacs

using namespace bacs;

extern "C" {

size_t bacs_distill_workspace_bytes(int B, int A, int h, int w, int H, int W) {
  if (B <= 0 || A <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return 0;
  size_t need = distill_layout(B, A, h, w, H, W).total;
  // the tensor-core path keeps boundary rows and loss partials per CTA; sized for the widest storage type
  need = std::max(need, distill_tc_workspace_bytes(BACS_F32, B, A, h, w, H, W));
  need = std::max(need, distill_tc_workspace_bytes(BACS_BF16, B, A, h, w, H, W));
  return need;
}

int bacs_distill_set_mode(int mode) {
  BACS_REQUIRE(mode >= 0 && mode <= 2, "bacs_distill_set_mode: mode must be 0 (auto), 1 (FMA kernel) or 2 (tensor cores)");
  g_distill_mode = mode;
  return BACS_OK;
}

int bacs_distill_kernel_variant(int dtype, int B, int A, int h, int w, int H, int W) {
  return (g_distill_mode != 1 && distill_tc_shape_ok(dtype, B, A, h, w, H, W)) ? 1 : 0;
}

static int teacher_distill_impl(const void* old_att, const void* new_att, int dtype, int B, int A, int h, int w,
                                const uint8_t* mask, int H, int W, float grad_coef, double* loss_sum, float* loss_scaled,
                                const float* addend, void* dnew, void* workspace, size_t workspace_bytes,
                                bacs_stream_t stream) {
  BACS_REQUIRE(old_att && new_att && loss_sum && workspace, "bacs_teacher_distill: null pointer");
  BACS_REQUIRE(B > 0 && B < 65536 && A > 0 && h > 0 && w > 0, "bacs_teacher_distill: bad shape");
  BACS_REQUIRE(H >= h && W >= w, "bacs_teacher_distill: the mask must be at least as large as the attention map");
  if (g_distill_mode != 1) {
    const int rc = distill_tc_launch(old_att, new_att, dtype, B, A, h, w, mask, H, W, grad_coef, loss_sum, loss_scaled, addend, dnew,
                                     workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc <= 0) return rc;  // done, or a real error
    if (g_distill_mode == 2) {
      set_error("bacs_teacher_distill: the tensor-core path does not apply to this shape / alignment");
      return BACS_ERR_UNSUPPORTED;
    }
  }
  if (w > 128) {
    set_error("bacs_teacher_distill: attention width %d > 128 not supported", w);
    return BACS_ERR_UNSUPPORTED;
  }
  const DistillLayout l = distill_layout(B, A, h, w, H, W);
  if (workspace_bytes < l.total) {
    set_error("bacs_teacher_distill: workspace %zu < %zu", workspace_bytes, l.total);
    return BACS_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  DistillTables tb;
  tb.yi0 = reinterpret_cast<int*>(ws + l.off_yi0);
  tb.ywt = reinterpret_cast<float*>(ws + l.off_ywt);
  tb.xi0 = reinterpret_cast<int*>(ws + l.off_xi0);
  tb.xwt = reinterpret_cast<float*>(ws + l.off_xwt);
  tb.rowstart = reinterpret_cast<int*>(ws + l.off_rowstart);
  tb.colstart = reinterpret_cast<int*>(ws + l.off_colstart);
  float* moments = reinterpret_cast<float*>(ws + l.off_moments);
  double* partials = reinterpret_cast<double*>(ws + l.off_partials);
  cudaStream_t s = (cudaStream_t)stream;

  distill_tables_kernel<<<1, 256, 0, s>>>(tb, H, W, h, w, hp_scale(h, H), hp_scale(w, W));
  BACS_CHECK_LAUNCH("bacs_teacher_distill(tables)");
  const int64_t nm = (int64_t)B * H * w;
  distill_moments_kernel<<<(unsigned)((nm + 255) / 256), 256, 0, s>>>(mask, B, H, W, w, tb.colstart, tb.xwt, moments);
  BACS_CHECK_LAUNCH("bacs_teacher_distill(moments)");

  // an interval holds the rows whose source row is i: at most ceil(H/h) + ceil(H/(2h)) + 2
  const int rows_max = (H + h - 1) / h + (H + 2 * h - 1) / (2 * h) + 2;
  const int wp = w <= 32 ? 32 : (w <= 64 ? 64 : 128);
  const size_t smem = sizeof(float) * ((size_t)3 * rows_max * 5 * wp + H + h + 4);
  if (smem > 200 * 1024) {
    set_error("bacs_teacher_distill: up-sampling ratio too large for the shared-memory moment tile");
    return BACS_ERR_UNSUPPORTED;
  }
  const int64_t total_units = (int64_t)B * l.n_cta_x * h;
  if (total_units > 0x3fffffff) {
    set_error("bacs_teacher_distill: problem too large");
    return BACS_ERR_UNSUPPORTED;
  }
  const int grid = (int)std::min<int64_t>(total_units, 2 * sm_count());
#define LAUNCH_DISTILL(TT, CPL)                                                                                   \
  do {                                                                                                            \
    auto kern = distill_kernel<TT, CPL>;                                                                          \
    if (smem > 48 * 1024) {                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
      if (e != cudaSuccess) {                                                                                     \
        set_error("bacs_teacher_distill: shared memory opt-in failed: %s", cudaGetErrorString(e));               \
        return BACS_ERR_CUDA;                                                                                     \
      }                                                                                                           \
    }                                                                                                             \
    kern<<<grid, 32 * kDistillWarps, smem, s>>>(reinterpret_cast<const TT*>(old_att),                            \
                                                reinterpret_cast<const TT*>(new_att), A, h, w, H, tb, moments,   \
                                                rows_max, grad_coef, reinterpret_cast<TT*>(dnew), partials,      \
                                                l.n_cta_x, (int)total_units);                                     \
  } while (0)
  BACS_DISPATCH_DTYPE(dtype, TT, {
    if (w <= 32) LAUNCH_DISTILL(TT, 1);
    else if (w <= 64) LAUNCH_DISTILL(TT, 2);
    else LAUNCH_DISTILL(TT, 4);
  });
#undef LAUNCH_DISTILL
  BACS_CHECK_LAUNCH("bacs_teacher_distill");
  sum_partials_kernel<<<1, 1024, 0, s>>>(partials, grid, loss_sum, grad_coef, loss_scaled, addend);
  BACS_CHECK_LAUNCH("bacs_teacher_distill(reduce)");
  return BACS_OK;
}

int bacs_teacher_distill(const void* old_att, const void* new_att, int dtype, int B, int A, int h, int w,
                         const uint8_t* mask, int H, int W, float grad_coef, double* loss_sum, float* loss_scaled, void* dnew,
                         void* workspace, size_t workspace_bytes, bacs_stream_t stream) {
  return teacher_distill_impl(old_att, new_att, dtype, B, A, h, w, mask, H, W, grad_coef, loss_sum, loss_scaled, nullptr, dnew,
                              workspace, workspace_bytes, stream);
}

int bacs_teacher_distill_add(const void* old_att, const void* new_att, int dtype, int B, int A, int h, int w,
                             const uint8_t* mask, int H, int W, float grad_coef, double* loss_sum, float* loss_scaled,
                             const float* addend, void* dnew, void* workspace, size_t workspace_bytes,
                             bacs_stream_t stream) {
  BACS_REQUIRE(loss_scaled && addend, "bacs_teacher_distill_add: null pointer");
  return teacher_distill_impl(old_att, new_att, dtype, B, A, h, w, mask, H, W, grad_coef, loss_sum, loss_scaled, addend, dnew,
                              workspace, workspace_bytes, stream);
}

}  // extern "C"
