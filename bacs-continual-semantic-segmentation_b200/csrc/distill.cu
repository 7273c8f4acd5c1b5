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
  for (int X = x0; X < x1; ++X) {
    if (mrow == nullptr || mrow[X]) {
      const float t = xwt[X] - 0.5f;
      const float t2 = t * t;
      m0 += 1.f;
      m1 += t;
      m2 += t2;
      m3 += t2 * t;
      m4 += t2 * t2;
    }
  }
  float* out = moments + row * 5 * w + j;
  out[0] = m0;
  out[w] = m1;
  out[2 * w] = m2;
  out[3 * w] = m3;
  out[4 * w] = m4;
}

constexpr int kDistillWarps = 8;
constexpr int kChanPerWarp = 2;
constexpr int kChanPerCta = kDistillWarps * kChanPerWarp;

template <typename T, int CPL>
__global__ void __launch_bounds__(32 * kDistillWarps) distill_kernel(const T* __restrict__ old_att,
                                                                      const T* __restrict__ new_att, int A, int h,
                                                                      int w, int H, DistillTables tb,
                                                                      const float* __restrict__ moments, int rows_max,
                                                                      float grad_coef, T* __restrict__ dnew,
                                                                      double* __restrict__ partials) {
  extern __shared__ float s_mom[];  // [rows_max][5][w]
  __shared__ double red_scratch[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int ch0 = blockIdx.x * kChanPerCta + wid * kChanPerWarp;
  const bool want_grad = dnew != nullptr;

  // per-lane cells: j = lane + 32*m
  float o_prev[kChanPerWarp][CPL], o_prev1[kChanPerWarp][CPL], n_prev[kChanPerWarp][CPL], n_prev1[kChanPerWarp][CPL];
  float carry[kChanPerWarp][CPL];
  float loss_acc = 0.f;

  auto load_row = [&](const T* base, int ch, int row, float* v, float* v1) {
#pragma unroll
    for (int m = 0; m < CPL; ++m) {
      const int j = lane + 32 * m;
      v[m] = v1[m] = 0.f;
      if (j < w && ch < A) {
        const T* p = base + (((int64_t)b * A + ch) * h + row) * w;
        v[m] = DT<T>::to_f(p[j]);
        v1[m] = DT<T>::to_f(p[min(j + 1, w - 1)]);
      }
    }
  };
#pragma unroll
  for (int c = 0; c < kChanPerWarp; ++c) {
    load_row(old_att, ch0 + c, 0, o_prev[c], o_prev1[c]);
    load_row(new_att, ch0 + c, 0, n_prev[c], n_prev1[c]);
#pragma unroll
    for (int m = 0; m < CPL; ++m) carry[c][m] = 0.f;
  }

  for (int i = 0; i < h; ++i) {
    const int Y0 = tb.rowstart[i], Y1 = tb.rowstart[i + 1];
    const int nrows = Y1 - Y0;
    const int row1 = min(i + 1, h - 1);
    // stage the moments of this interval's rows (shared by every channel of the CTA)
    __syncthreads();
    {
      const float* src = moments + ((int64_t)b * H + Y0) * 5 * w;
      const int n = nrows * 5 * w;
      for (int k = threadIdx.x; k < n; k += blockDim.x) s_mom[k] = __ldg(src + k);
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kChanPerWarp; ++c) {
      const int ch = ch0 + c;
      float o_cur[CPL], o_cur1[CPL], n_cur[CPL], n_cur1[CPL];
      if (row1 != i) {
        load_row(old_att, ch, row1, o_cur, o_cur1);
        load_row(new_att, ch, row1, n_cur, n_cur1);
      } else {
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          o_cur[m] = o_prev[c][m]; o_cur1[m] = o_prev1[c][m];
          n_cur[m] = n_prev[c][m]; n_cur1[m] = n_prev1[c][m];
        }
      }
      // gradient accumulators w.r.t. new[row i][j], new[row1][j] and the same at column j+1
      float g0[CPL], g1[CPL], g0n[CPL], g1n[CPL];
#pragma unroll
      for (int m = 0; m < CPL; ++m) g0[m] = g1[m] = g0n[m] = g1n[m] = 0.f;

      for (int r = 0; r < nrows; ++r) {
        const float ty = tb.ywt[Y0 + r];
        const float sy0 = 1.f - ty;
        const float* mr = s_mom + r * 5 * w;
        float a_n[CPL], d_n[CPL], c0[CPL], c1[CPL], c2[CPL];
        float M0[CPL], M1[CPL], M2[CPL], M3[CPL], M4[CPL];
        float S = 0.f;
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const int j = lane + 32 * m;
          const bool on = j < w;
          const float ro = sy0 * o_prev[c][m] + ty * o_cur[m];
          const float ro1 = sy0 * o_prev1[c][m] + ty * o_cur1[m];
          const float rn = sy0 * n_prev[c][m] + ty * n_cur[m];
          const float rn1 = sy0 * n_prev1[c][m] + ty * n_cur1[m];
          const float d_o = ro1 - ro;
          const float a_o = ro + 0.5f * d_o;
          d_n[m] = rn1 - rn;
          a_n[m] = rn + 0.5f * d_n[m];
          c0[m] = (a_o - a_n[m]) * (a_o + a_n[m]);
          // a_o d_o - a_n d_n written so that identical maps give exactly 0 (no FMA residue)
          c1[m] = 2.f * ((a_o - a_n[m]) * d_o + a_n[m] * (d_o - d_n[m]));
          c2[m] = (d_o - d_n[m]) * (d_o + d_n[m]);
          M0[m] = on ? mr[j] : 0.f;
          M1[m] = on ? mr[w + j] : 0.f;
          M2[m] = on ? mr[2 * w + j] : 0.f;
          M3[m] = on ? mr[3 * w + j] : 0.f;
          M4[m] = on ? mr[4 * w + j] : 0.f;
          S += c0[m] * c0[m] * M0[m] + 2.f * c0[m] * c1[m] * M1[m] + (c1[m] * c1[m] + 2.f * c0[m] * c2[m]) * M2[m] +
               2.f * c1[m] * c2[m] * M3[m] + c2[m] * c2[m] * M4[m];
        }
        S = warp_sum(S);
        if (!(S > 0.f)) continue;  // zero (or rounding-negative) row: norm 0, sub-gradient 0
        const float nrm = sqrtf(S);
        loss_acc += nrm;
        if (!want_grad) continue;
        const float gS = 0.5f / nrm;
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const float dS0 = 2.f * (c0[m] * M0[m] + c1[m] * M1[m] + c2[m] * M2[m]);
          const float dS1 = 2.f * (c0[m] * M1[m] + c1[m] * M2[m] + c2[m] * M3[m]);
          const float dS2 = 2.f * (c0[m] * M2[m] + c1[m] * M3[m] + c2[m] * M4[m]);
          const float gA = gS * (-2.f * a_n[m] * dS0 - 2.f * d_n[m] * dS1);
          const float gD = gS * (-2.f * a_n[m] * dS1 - 2.f * d_n[m] * dS2);
          const float gj = 0.5f * gA - gD;   // d/d rn[j]
          const float gj1 = 0.5f * gA + gD;  // d/d rn[j+1]
          g0[m] += sy0 * gj;
          g1[m] += ty * gj;
          g0n[m] += sy0 * gj1;
          g1n[m] += ty * gj1;
        }
      }

      if (want_grad) {
        // move the column-(j+1) contributions to their owner (last column owns its own)
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const int j = lane + 32 * m;
          // lane 0 of group m takes the value of lane 31 of group m-1 (all lanes run both shuffles)
          const float wrap0 = __shfl_sync(0xffffffffu, m > 0 ? g0n[m > 0 ? m - 1 : 0] : 0.f, 31);
          const float wrap1 = __shfl_sync(0xffffffffu, m > 0 ? g1n[m > 0 ? m - 1 : 0] : 0.f, 31);
          const float sh0 = __shfl_up_sync(0xffffffffu, g0n[m], 1);
          const float sh1 = __shfl_up_sync(0xffffffffu, g1n[m], 1);
          const float up0 = lane == 0 ? wrap0 : sh0;
          const float up1 = lane == 0 ? wrap1 : sh1;
          if (j > 0 && j < w) {
            g0[m] += up0;
            g1[m] += up1;
          }
          if (j == w - 1) {
            g0[m] += g0n[m];
            g1[m] += g1n[m];
          }
        }
        // row i is complete: carry from the interval above + this interval's share
#pragma unroll
        for (int m = 0; m < CPL; ++m) {
          const int j = lane + 32 * m;
          float fin = carry[c][m] + g0[m];
          if (row1 == i) fin += g1[m];
          if (j < w && ch < A) dnew[(((int64_t)b * A + ch) * h + i) * w + j] = DT<T>::from_f(grad_coef * fin);
          carry[c][m] = g1[m];
        }
      }
#pragma unroll
      for (int m = 0; m < CPL; ++m) {
        o_prev[c][m] = o_cur[m]; o_prev1[c][m] = o_cur1[m];
        n_prev[c][m] = n_cur[m]; n_prev1[c][m] = n_cur1[m];
      }
    }
  }
  // every lane holds the same loss_acc (warp_sum broadcasts); count it once per warp
  const double mine = (lane == 0 && ch0 < A) ? (double)loss_acc : 0.0;
  const double tot = block_sum(mine, red_scratch);
  if (threadIdx.x == 0) partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(const double* __restrict__ partials, int n,
                                                            double* __restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
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
  l.n_cta_x = (A + kChanPerCta - 1) / kChanPerCta;
  l.off_partials = take(sizeof(double) * (size_t)B * l.n_cta_x);
  l.total = o;
  return l;
}

}  // namespace bacs

using namespace bacs;

extern "C" {

size_t bacs_distill_workspace_bytes(int B, int A, int h, int w, int H, int W) {
  if (B <= 0 || A <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return 0;
  return distill_layout(B, A, h, w, H, W).total;
}

int bacs_teacher_distill(const void* old_att, const void* new_att, int dtype, int B, int A, int h, int w,
                         const uint8_t* mask, int H, int W, float grad_coef, double* loss_sum, void* dnew,
                         void* workspace, size_t workspace_bytes, bacs_stream_t stream) {
  BACS_REQUIRE(old_att && new_att && loss_sum && workspace, "bacs_teacher_distill: null pointer");
  BACS_REQUIRE(B > 0 && B < 65536 && A > 0 && h > 0 && w > 0, "bacs_teacher_distill: bad shape");
  BACS_REQUIRE(H >= h && W >= w, "bacs_teacher_distill: the mask must be at least as large as the attention map");
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
  const size_t smem = sizeof(float) * (size_t)rows_max * 5 * w;
  if (smem > 200 * 1024) {
    set_error("bacs_teacher_distill: up-sampling ratio too large for the shared-memory moment tile");
    return BACS_ERR_UNSUPPORTED;
  }
  dim3 grid(l.n_cta_x, B);
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
                                                rows_max, grad_coef, reinterpret_cast<TT*>(dnew), partials);     \
  } while (0)
  BACS_DISPATCH_DTYPE(dtype, TT, {
    if (w <= 32) LAUNCH_DISTILL(TT, 1);
    else if (w <= 64) LAUNCH_DISTILL(TT, 2);
    else LAUNCH_DISTILL(TT, 4);
  });
#undef LAUNCH_DISTILL
  BACS_CHECK_LAUNCH("bacs_teacher_distill");
  sum_partials_kernel<<<1, 1024, 0, s>>>(partials, B * l.n_cta_x, loss_sum);
  BACS_CHECK_LAUNCH("bacs_teacher_distill(reduce)");
  return BACS_OK;
}

}  // extern "C"
