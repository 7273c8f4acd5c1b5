// DER logit MSE, confusion matrix + metrics, device-side scalar helpers, state packing.
#include <algorithm>
#include "common.cuh"

namespace bacs {

// -------------------------------------------------------------------------------------
// Dark-experience-replay logit MSE with transplant (loss/bacs_loss.py:387-431)
// -------------------------------------------------------------------------------------
template <typename T, typename M>
__global__ void __launch_bounds__(256) der_mse_kernel(const T* __restrict__ sem, const M* __restrict__ mem,
                                                      int truncate, const int32_t* __restrict__ cut, int ignore_rep_bg,
                                                      int K, int hw, int64_t total, float grad_coef,
                                                      T* __restrict__ dsem, double* __restrict__ partials) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t jk = i / hw;
    const int k = (int)(jk % K);
    const int j = (int)(jk / K);
    const float s = DT<T>::to_f(sem[i]);
    float m;
    if (k >= cut[j] || (ignore_rep_bg && k == 0)) {
      m = s;  // transplanted from the live logits
    } else {
      m = (float)mem[i];
      if (truncate) m = truncf(m);  // preprocess_batch's .long() (base_loss.py:274-278)
    }
    const float d = s - m;
    acc = fmaf(d, d, acc);
    if (dsem) dsem[i] = DT<T>::from_f(2.f * grad_coef * d);
  }
  const float r = block_sum(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = (double)r;
}

__global__ void __launch_bounds__(1024) sum_partials2_kernel(const double* __restrict__ partials, int n,
                                                             double* __restrict__ out) {
  __shared__ double scratch[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s;
}

// Transplant cut per replay sample with the reference's index quirk (bacs_loss.py:415-425):
//   u, inv = unique(n_classes, return_inverse=True); for i, n in enumerate(u): j = inv[i];
//   if n < K: memory[j, n:] = live[j, n:]      -- j is ONE sample index (Q5), not a mask.
// Single block; values are class counts in [0, 1024).
__global__ void __launch_bounds__(256) der_cut_kernel(const int64_t* __restrict__ n_classes, int Br, int K,
                                                      int32_t* __restrict__ cut) {
  __shared__ int present[1024];
  __shared__ int rank_of[1024];
  __shared__ int uniq[1024];
  __shared__ int warp_tot[8];
  __shared__ int s_first[1024];  // class count (clamped) of the first min(Br, 1024) samples
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  auto clampv = [](int64_t n) { return n < 0 ? 0 : (n > 1023 ? 1023 : (int)n); };
  for (int i = tid; i < 1024; i += blockDim.x) present[i] = 0;
  __syncthreads();
  for (int j = tid; j < Br; j += blockDim.x) {
    const int v = clampv(n_classes[j]);
    present[v] = 1;
    if (j < 1024) s_first[j] = v;
    cut[j] = K;
  }
  __syncthreads();
  // rank of every present value = exclusive prefix count: thread t owns values 4t .. 4t+3 (block scan)
  int p[4], mine = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    p[e] = present[4 * tid + e];
    mine += p[e];
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  int base = incl - mine;
  for (int wv = 0; wv < wid; ++wv) base += warp_tot[wv];
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (p[e]) {
      rank_of[4 * tid + e] = base;
      uniq[base] = 4 * tid + e;
      ++base;
    }
  __syncthreads();
  if (tid == 0) {
    int r = 0;
    for (int wv = 0; wv < 8; ++wv) r += warp_tot[wv];
    // sequential like the reference's Python loop (later writes win, all are mins of the same slot);
    // i < r <= 1024 distinct values and i < Br, so s_first covers every sample the loop reads
    for (int i = 0; i < r && i < Br; ++i) {
      const int j = rank_of[s_first[i]];  // inv[i]
      if (uniq[i] < K && j < Br) cut[j] = min(cut[j], uniq[i]);
    }
  }
}

// -------------------------------------------------------------------------------------
// MiB unbiased knowledge distillation (training/loss_utils.py:447-489); MiB / SDR only.
//   per = ( q_0 (lse_{0 u new}(x) - lse(x)) + sum_{c=1}^{Ko-1} q_c (x_c - lse(x)) ) / Ko,  q = softmax(alpha * old)
//   loss = -mean(mask * per);  d per / d x_k = ( q_0 [k in {0} u new] e^{x_k} / S_B + [1 <= k < Ko] q_k - P_k ) / Ko
// One thread per pixel; channel planes are read with unit stride across the warp.
// -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) unbiased_kd_kernel(const T* __restrict__ x, const T* __restrict__ old, int K,
                                                          int Ko, int64_t HW, int64_t npix, float alpha,
                                                          const uint8_t* __restrict__ mask, float grad_coef,
                                                          T* __restrict__ dx, double* __restrict__ partials) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, q = p - b * HW;
    const T* xp = x + b * K * HW + q;
    const T* op = old + b * Ko * HW + q;
    float mx = -INFINITY, mo = -INFINITY;
    for (int c = 0; c < K; ++c) mx = fmaxf(mx, DT<T>::to_f(xp[(int64_t)c * HW]));
    for (int c = 0; c < Ko; ++c) mo = fmaxf(mo, alpha * DT<T>::to_f(op[(int64_t)c * HW]));
    float S = 0.f, SB = 0.f, So = 0.f;
    for (int c = 0; c < K; ++c) {
      const float e = __expf(DT<T>::to_f(xp[(int64_t)c * HW]) - mx);
      S += e;
      if (c == 0 || c >= Ko) SB += e;
    }
    for (int c = 0; c < Ko; ++c) So += __expf(alpha * DT<T>::to_f(op[(int64_t)c * HW]) - mo);
    const float den = mx + __logf(S);
    const float inv_So = 1.f / So;
    const float q0 = __expf(alpha * DT<T>::to_f(op[0]) - mo) * inv_So;
    float per = q0 * (mx + __logf(SB) - den);
    for (int c = 1; c < Ko; ++c) {
      const float qc = __expf(alpha * DT<T>::to_f(op[(int64_t)c * HW]) - mo) * inv_So;
      per = fmaf(qc, DT<T>::to_f(xp[(int64_t)c * HW]) - den, per);
    }
    const float m = mask ? (mask[p] ? 1.f : 0.f) : 1.f;
    acc += m * per / (float)Ko;
    if (dx) {
      T* dp = dx + b * K * HW + q;
      const float g = -grad_coef * m / (float)Ko;
      const float inv_S = 1.f / S, inv_SB = 1.f / SB;
      for (int c = 0; c < K; ++c) {
        const float e = __expf(DT<T>::to_f(xp[(int64_t)c * HW]) - mx);
        float d = -e * inv_S;
        if (c == 0 || c >= Ko) d += q0 * e * inv_SB;
        else d += __expf(alpha * DT<T>::to_f(op[(int64_t)c * HW]) - mo) * inv_So;
        dp[(int64_t)c * HW] = DT<T>::from_f(g * d);
      }
    }
  }
  const float r = block_sum(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = (double)r;
}

static int der_blocks(int64_t total) {
  int64_t b = (total + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = (int64_t)sm_count() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// -------------------------------------------------------------------------------------
// Confusion matrix: block-private K*K histogram in shared memory, run-length aggregated
// per thread, flushed with 64-bit global atomics.
//   reference: training/metrics.py:38-50 + torchmetrics 0.6.0 _confusion_matrix_update
// -------------------------------------------------------------------------------------
template <typename P, bool SMEM>
__global__ void __launch_bounds__(256) confmat_kernel(const P* __restrict__ preds, const int64_t* __restrict__ target,
                                                      int64_t n, int K, unsigned long long* __restrict__ confmat,
                                                      unsigned long long* __restrict__ oob) {
  extern __shared__ unsigned int sh[];
  const int KK = K * K;
  if (SMEM) {
    for (int i = threadIdx.x; i < KK; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  unsigned int n_oob = 0;
  const int64_t nchunk = (n + 7) >> 3;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t base = c << 3;
    const int cnt = (int)min((int64_t)8, n - base);
    int cur = -1;
    unsigned int run = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < cnt) {
        // metrics.py:45-46 casts both to int32 before the range test
        const int t = (int)target[base + k];
        const int p = (int)preds[base + k];
        int bin = -1;
        if (t >= 0 && t < K) {
          if (p >= 0 && p < K) bin = t * K + p;
          else ++n_oob;
        }
        if (bin == cur) {
          ++run;
        } else {
          if (run && cur >= 0) {
            if (SMEM) atomicAdd(&sh[cur], run);
            else atomicAdd(&confmat[cur], (unsigned long long)run);
          }
          cur = bin;
          run = 1;
        }
      }
    }
    if (run && cur >= 0) {
      if (SMEM) atomicAdd(&sh[cur], run);
      else atomicAdd(&confmat[cur], (unsigned long long)run);
    }
  }
  if (n_oob && oob) atomicAdd(oob, (unsigned long long)n_oob);
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < KK; i += blockDim.x)
      if (sh[i]) atomicAdd(&confmat[i], (unsigned long long)sh[i]);
  }
}


// int64 predictions, 16-byte aligned: coalesced 16-byte loads (a lane owns two adjacent pixels per load, four loads of
// each tensor in flight) and warp-level aggregation -- a warp inside a uniform region (the usual case in a
// segmentation map) issues ONE shared-memory atomic per pixel slot: the lanes that agree with the first valid lane
// are counted by a ballot, only the others (region borders, wrong pixels) add on their own.
__device__ __forceinline__ void confmat_warp_add(unsigned int* sh, int bin, int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  const unsigned valid = __ballot_sync(kFull, bin >= 0);
  if (valid == 0) return;
  const int leader = __ffs(valid) - 1;
  const int first = __shfl_sync(kFull, bin, leader);
  const unsigned same = __ballot_sync(kFull, bin == first);   // the leader's bin: one atomic for all its lanes
  if (lane == leader) atomicAdd(&sh[first], (unsigned int)__popc(same));
  else if (bin >= 0 && bin != first) atomicAdd(&sh[bin], 1u);  // the others (region borders, wrong pixels) on their own
}

__global__ void __launch_bounds__(256) confmat_vec_kernel(const longlong2* __restrict__ preds,
                                                          const longlong2* __restrict__ target, int64_t nvec, int K,
                                                          unsigned long long* __restrict__ confmat,
                                                          unsigned long long* __restrict__ oob) {
  extern __shared__ unsigned int sh[];
  const int KK = K * K, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < KK; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  unsigned int n_oob = 0;
  constexpr int U = 4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x * U; v0 < nvec; v0 += step) {  // warp-uniform trip count
    longlong2 t[U], p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + (int64_t)u * blockDim.x + threadIdx.x;
      if (v < nvec) {
        t[u] = __ldg(target + v);
        p[u] = __ldg(preds + v);
      } else {
        t[u] = make_longlong2(-1, -1);
        p[u] = make_longlong2(0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // metrics.py:45-46 casts both to int32 before the range test
      const int t0 = (int)t[u].x, t1 = (int)t[u].y, p0 = (int)p[u].x, p1 = (int)p[u].y;
      int b0 = -1, b1 = -1;
      if (t0 >= 0 && t0 < K) {
        if (p0 >= 0 && p0 < K) b0 = t0 * K + p0;
        else ++n_oob;
      }
      if (t1 >= 0 && t1 < K) {
        if (p1 >= 0 && p1 < K) b1 = t1 * K + p1;
        else ++n_oob;
      }
      confmat_warp_add(sh, b0, lane);
      confmat_warp_add(sh, b1, lane);
    }
  }
  if (n_oob && oob) atomicAdd(oob, (unsigned long long)n_oob);
  __syncthreads();
  for (int i = threadIdx.x; i < KK; i += blockDim.x)
    if (sh[i]) atomicAdd(&confmat[i], (unsigned long long)sh[i]);
}

// metrics.py:52-88 (the reference's fp/fn naming is kept: "fn" = column sum - tp,
// "fp" = row sum - tp) + torchmetrics' IoU from the matrix (absent -> 0, reduction none).
__global__ void __launch_bounds__(256) confmat_metrics_kernel(const int64_t* __restrict__ cm, int K,
                                                              float* __restrict__ out) {
  __shared__ float scratch[32];
  __shared__ long long s_total;
  __shared__ float s_miou;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  long long part = 0;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) part += cm[i];
  // exact integer block sum
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)&s_total, (unsigned long long)part);
  __syncthreads();
  const long long total = s_total;
  float iou_sum = 0.f;
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    long long row = 0, col = 0;
    for (int k = 0; k < K; ++k) {
      row += cm[c * K + k];
      col += cm[k * K + c];
    }
    const long long tp = cm[c * K + c];
    const long long fn = col - tp, fp = row - tp;
    const long long tn = total - (tp + fn + fp);
    auto div = [](long long a, long long b) {
      const float r = (float)a / (float)b;
      return isnan(r) ? 0.f : r;
    };
    const long long uni = row + col - tp;
    const float iou = uni == 0 ? 0.f : (float)tp / (float)uni;
    out[0 * K + c] = iou;
    out[1 * K + c] = div(tp + tn, tp + fp + fn + tn);
    out[2 * K + c] = div(tp, tp + fp);
    out[3 * K + c] = div(tp, tp + fn);
    out[4 * K + c] = div(tn, tn + fp);
    iou_sum += iou;
  }
  const float s = block_sum(iou_sum, scratch);
  if (threadIdx.x == 0) s_miou = s / (float)K;
  __syncthreads();
  for (int c = threadIdx.x; c < K; c += blockDim.x) out[5 * K + c] = s_miou;
}

// -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ g) {
  const float s = *g;
  if (s == 1.f) return;  // uniform early exit: the common (bf16, no grad scaler) case costs nothing
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = DT<T>::from_f(DT<T>::to_f(x[i]) * s);
}

// Up to 8 tensors in ONE launch (the backward of the fused per-pixel Function rescales logits-, feature-, weight-
// and bias-gradients by the same upstream scalar): block ranges are assigned per tensor.
struct ScaleMulti {
  void* p[8];
  int64_t n[8];
  int dt[8];
  int blk0[9];
  int cnt;
};
template <typename T>
__device__ __forceinline__ void scale_range(T* x, int64_t n, float s, int lb, int nb) {
  for (int64_t i = (int64_t)lb * blockDim.x + threadIdx.x; i < n; i += (int64_t)nb * blockDim.x)
    x[i] = DT<T>::from_f(DT<T>::to_f(x[i]) * s);
}
__global__ void __launch_bounds__(256) scale_multi_kernel(const ScaleMulti a, const float* __restrict__ g) {
  const float s = *g;
  if (s == 1.f) return;
  int i = 0;
  while (i + 1 < a.cnt && (int)blockIdx.x >= a.blk0[i + 1]) ++i;
  const int lb = (int)blockIdx.x - a.blk0[i], nb = a.blk0[i + 1] - a.blk0[i];
  if (a.dt[i] == BACS_F32) scale_range(reinterpret_cast<float*>(a.p[i]), a.n[i], s, lb, nb);
  else if (a.dt[i] == BACS_BF16) scale_range(reinterpret_cast<__nv_bfloat16*>(a.p[i]), a.n[i], s, lb, nb);
  else scale_range(reinterpret_cast<__half*>(a.p[i]), a.n[i], s, lb, nb);
}

__global__ void pack_state_kernel(const double* __restrict__ sums, const double* __restrict__ counts, int TD, int T,
                                  const int64_t* __restrict__ confmat, int KK, double* __restrict__ packed) {
  const int n = TD + T + KK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v;
    if (i < TD) v = sums[i];
    else if (i < TD + T) v = counts[i - TD];
    else v = (double)confmat[i - TD - T];
    packed[i] = v;
  }
}
__global__ void unpack_state_kernel(const double* __restrict__ packed, int TD, int T, double* __restrict__ sums,
                                    double* __restrict__ counts, int64_t* __restrict__ confmat, int KK) {
  const int n = TD + T + KK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = packed[i];
    if (i < TD) sums[i] = v;
    else if (i < TD + T) counts[i - TD] = v;
    else confmat[i - TD - T] = (int64_t)llrint(v);
  }
}

// loss = sum_i coef[i] * src_i[idx_i] * (gate ? [*gate != 0] : 1) / (den_i ? den_i[didx_i] : 1)
struct CombineTerm {
  const double* src;
  int idx;
  const double* den;
  int didx;
  float coef;
};
struct CombineArgs {
  CombineTerm t[8];
  int n;
};
__global__ void combine_kernel(CombineArgs a, float* __restrict__ out) {
  double s = 0.0;
  for (int i = 0; i < a.n; ++i) {
    double v = a.t[i].src[a.t[i].idx] * (double)a.t[i].coef;
    if (a.t[i].den) {
      v = v / a.t[i].den[a.t[i].didx];  // IEEE: 0/0 -> NaN exactly like torch's mean over nothing
    }
    s += v;
  }
  out[0] = (float)s;
}

}  // namespace bacs

using namespace bacs;

namespace bacs {
// dst[i] = src[idx[i]] for rows of `row_elems` V's: the replay store's minibatch gather (an index outside [0, n_rows)
// yields a zero row instead of a fault)
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ src, int64_t n_rows, int64_t row_elems,
                                                          const int64_t* __restrict__ idx, V* __restrict__ dst) {
  const int64_t r = idx[blockIdx.y];
  const bool ok = r >= 0 && r < n_rows;
  const V* s = src + (ok ? r : 0) * row_elems;
  V* d = dst + (int64_t)blockIdx.y * row_elems;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < row_elems; i += stride) d[i] = ok ? s[i] : V{};
}
}  // namespace bacs

extern "C" {

size_t bacs_der_workspace_bytes(int Br, int K, int hw) {
  return align_up(sizeof(double) * (size_t)der_blocks((int64_t)Br * K * hw), 256);
}

int bacs_der_mse(const void* sem_logits, int dtype, const void* memory_logits, int memory_is_int64, int truncate,
                 const int32_t* cut, int ignore_rep_bg, int Br, int K, int hw, float grad_coef, double* loss_sum,
                 void* dsem, void* workspace, size_t workspace_bytes, bacs_stream_t stream) {
  BACS_REQUIRE(sem_logits && memory_logits && cut && loss_sum && workspace, "bacs_der_mse: null pointer");
  BACS_REQUIRE(Br > 0 && K > 0 && hw > 0, "bacs_der_mse: bad shape");
  const int64_t total = (int64_t)Br * K * hw;
  const int blocks = der_blocks(total);
  if (workspace_bytes < sizeof(double) * (size_t)blocks) {
    set_error("bacs_der_mse: workspace too small");
    return BACS_ERR_WORKSPACE;
  }
  double* partials = reinterpret_cast<double*>(workspace);
  cudaStream_t s = (cudaStream_t)stream;
  BACS_DISPATCH_DTYPE(dtype, TT, {
    if (memory_is_int64)
      der_mse_kernel<TT, int64_t><<<blocks, 256, 0, s>>>(reinterpret_cast<const TT*>(sem_logits),
                                                         reinterpret_cast<const int64_t*>(memory_logits), 0, cut,
                                                         ignore_rep_bg, K, hw, total, grad_coef,
                                                         reinterpret_cast<TT*>(dsem), partials);
    else
      der_mse_kernel<TT, float><<<blocks, 256, 0, s>>>(reinterpret_cast<const TT*>(sem_logits),
                                                       reinterpret_cast<const float*>(memory_logits), truncate, cut,
                                                       ignore_rep_bg, K, hw, total, grad_coef,
                                                       reinterpret_cast<TT*>(dsem), partials);
  });
  BACS_CHECK_LAUNCH("bacs_der_mse");
  sum_partials2_kernel<<<1, 1024, 0, s>>>(partials, blocks, loss_sum);
  BACS_CHECK_LAUNCH("bacs_der_mse(reduce)");
  return BACS_OK;
}

size_t bacs_unbiased_kd_workspace_bytes(int64_t npix) {
  return align_up(sizeof(double) * (size_t)der_blocks(npix * 4), 256);
}

int bacs_unbiased_kd(const void* logits, const void* old_logits, int dtype, int B, int K, int K_old, int H, int W,
                     float alpha, const uint8_t* mask, float grad_coef, double* loss_sum, void* dlogits,
                     void* workspace, size_t workspace_bytes, bacs_stream_t stream) {
  BACS_REQUIRE(logits && old_logits && loss_sum && workspace, "bacs_unbiased_kd: null pointer");
  BACS_REQUIRE(B > 0 && K > 0 && K_old > 0 && K_old <= K && H > 0 && W > 0, "bacs_unbiased_kd: bad shape");
  const int64_t HW = (int64_t)H * W, npix = HW * B;
  const int blocks = der_blocks(npix * 4);
  if (workspace_bytes < sizeof(double) * (size_t)blocks) {
    set_error("bacs_unbiased_kd: workspace too small");
    return BACS_ERR_WORKSPACE;
  }
  double* partials = reinterpret_cast<double*>(workspace);
  cudaStream_t s = (cudaStream_t)stream;
  BACS_DISPATCH_DTYPE(dtype, TT, {
    unbiased_kd_kernel<TT><<<blocks, 256, 0, s>>>(reinterpret_cast<const TT*>(logits),
                                                  reinterpret_cast<const TT*>(old_logits), K, K_old, HW, npix, alpha,
                                                  mask, grad_coef, reinterpret_cast<TT*>(dlogits), partials);
  });
  BACS_CHECK_LAUNCH("bacs_unbiased_kd");
  sum_partials2_kernel<<<1, 1024, 0, s>>>(partials, blocks, loss_sum);
  BACS_CHECK_LAUNCH("bacs_unbiased_kd(reduce)");
  return BACS_OK;
}

int bacs_der_cut(const int64_t* n_classes, int Br, int K, int32_t* cut, bacs_stream_t stream) {
  BACS_REQUIRE(n_classes && cut && Br > 0 && K > 0, "bacs_der_cut: bad arguments");
  der_cut_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(n_classes, Br, K, cut);
  BACS_CHECK_LAUNCH("bacs_der_cut");
  return BACS_OK;
}

int bacs_confmat_accumulate(const void* preds, int preds_is_float, const int64_t* target, int64_t n, int K,
                            int64_t* confmat, int64_t* oob, bacs_stream_t stream) {
  BACS_REQUIRE(preds && target && confmat, "bacs_confmat_accumulate: null pointer");
  BACS_REQUIRE(K > 0 && K <= 4096 && n >= 0, "bacs_confmat_accumulate: bad K or n");
  if (n == 0) return BACS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nchunk = (n + 7) >> 3;
  int64_t blocks = (nchunk + 255) / 256;
  const size_t smem = sizeof(unsigned int) * (size_t)K * K;
  const bool use_smem = smem <= 96 * 1024;
  const int64_t cap = (int64_t)sm_count() * (use_smem && smem > 24 * 1024 ? 2 : 8);
  if (blocks > cap) blocks = cap;
  unsigned long long* cm = reinterpret_cast<unsigned long long*>(confmat);
  unsigned long long* ob = reinterpret_cast<unsigned long long*>(oob);
  if (!preds_is_float && smem <= 48 * 1024 && n >= 2 && ((reinterpret_cast<uintptr_t>(preds) | reinterpret_cast<uintptr_t>(target)) & 15) == 0) {
    const int64_t nvec = n >> 1;
    int64_t vb = (nvec + 256 * 4 - 1) / (256 * 4);
    vb = std::min<int64_t>(vb, (int64_t)sm_count() * 8);
    confmat_vec_kernel<<<(unsigned)vb, 256, smem, s>>>(reinterpret_cast<const longlong2*>(preds),
                                                       reinterpret_cast<const longlong2*>(target), nvec, K, cm, ob);
    BACS_CHECK_LAUNCH("bacs_confmat_accumulate");
    if (n & 1) {  // the odd last pixel
      confmat_kernel<int64_t, false><<<1, 32, 0, s>>>(reinterpret_cast<const int64_t*>(preds) + (n - 1), target + (n - 1), 1, K, cm, ob);
      BACS_CHECK_LAUNCH("bacs_confmat_accumulate(tail)");
    }
    return BACS_OK;
  }
#define LAUNCH_CM(PT)                                                                                          \
  do {                                                                                                         \
    if (use_smem) {                                                                                            \
      auto kern = confmat_kernel<PT, true>;                                                                    \
      if (smem > 48 * 1024) {                                                                                  \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        if (e != cudaSuccess) {                                                                                \
          set_error("bacs_confmat_accumulate: shared memory opt-in failed: %s", cudaGetErrorString(e));       \
          return BACS_ERR_CUDA;                                                                                \
        }                                                                                                      \
      }                                                                                                        \
      kern<<<(unsigned)blocks, 256, smem, s>>>(reinterpret_cast<const PT*>(preds), target, n, K, cm, ob);      \
    } else {                                                                                                   \
      confmat_kernel<PT, false><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const PT*>(preds), target, n, K, \
                                                                 cm, ob);                                      \
    }                                                                                                          \
  } while (0)
  if (preds_is_float) LAUNCH_CM(float);
  else LAUNCH_CM(int64_t);
#undef LAUNCH_CM
  BACS_CHECK_LAUNCH("bacs_confmat_accumulate");
  return BACS_OK;
}

int bacs_confmat_metrics(const int64_t* confmat, int K, float* out, bacs_stream_t stream) {
  BACS_REQUIRE(confmat && out && K > 0, "bacs_confmat_metrics: bad arguments");
  confmat_metrics_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(confmat, K, out);
  BACS_CHECK_LAUNCH("bacs_confmat_metrics");
  return BACS_OK;
}

int bacs_gather_rows(const void* src, int64_t n_rows, int64_t row_bytes, const int64_t* idx, int64_t n_idx, void* dst,
                     bacs_stream_t stream) {
  BACS_REQUIRE(src && idx && dst && n_rows > 0 && row_bytes > 0 && n_idx >= 0, "bacs_gather_rows: bad arguments");
  BACS_REQUIRE(n_idx <= 65535, "bacs_gather_rows: at most 65535 rows per call (got %lld)", (long long)n_idx);
  if (n_idx == 0) return BACS_OK;
  const bool vec = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  // blocks per row: enough to fill the machine a few times over, rows are MBs (images) or KBs (logits)
  const int64_t chunk = 256 * 16 * 4;  // bytes a block moves per sweep with 16-byte vectors, 4 in flight per thread
  int64_t per_row = std::max<int64_t>(1, std::min<int64_t>((row_bytes + chunk - 1) / chunk, 64));
  const int64_t cap = (int64_t)sm_count() * 16;
  if (per_row * n_idx > cap) per_row = std::max<int64_t>(1, cap / n_idx);
  dim3 grid((unsigned)per_row, (unsigned)n_idx);
  if (vec)
    gather_rows_kernel<uint4><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), n_rows,
                                                                      row_bytes / 16, idx, reinterpret_cast<uint4*>(dst));
  else
    gather_rows_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned char*>(src), n_rows, row_bytes, idx, reinterpret_cast<unsigned char*>(dst));
  BACS_CHECK_LAUNCH("bacs_gather_rows");
  return BACS_OK;
}

int bacs_scale_inplace(void* x, int dtype, int64_t n, const float* g_dev, bacs_stream_t stream) {
  BACS_REQUIRE(x && g_dev && n >= 0, "bacs_scale_inplace: bad arguments");
  if (n == 0) return BACS_OK;
  int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  BACS_DISPATCH_DTYPE(dtype, TT,
                      { scale_inplace_kernel<TT><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<TT*>(x), n, g_dev); });
  BACS_CHECK_LAUNCH("bacs_scale_inplace");
  return BACS_OK;
}

int bacs_scale_inplace_multi(int n, void* const* x, const int* dtype, const int64_t* numel, const float* g_dev,
                             bacs_stream_t stream) {
  BACS_REQUIRE(n >= 0 && n <= 8 && (n == 0 || (x && dtype && numel)) && g_dev, "bacs_scale_inplace_multi: bad arguments");
  ScaleMulti a;
  a.cnt = 0;
  a.blk0[0] = 0;
  const int64_t cap = (int64_t)sm_count() * 8;
  for (int i = 0; i < n; ++i) {
    if (!x[i] || numel[i] <= 0) continue;
    BACS_REQUIRE(dtype[i] >= 0 && dtype[i] <= BACS_F16, "bacs_scale_inplace_multi: unknown dtype %d", dtype[i]);
    int64_t blocks = (numel[i] + 256 * 8 - 1) / (256 * 8);
    if (blocks > cap) blocks = cap;
    a.p[a.cnt] = x[i];
    a.n[a.cnt] = numel[i];
    a.dt[a.cnt] = dtype[i];
    a.blk0[a.cnt + 1] = a.blk0[a.cnt] + (int)blocks;
    ++a.cnt;
  }
  if (a.cnt == 0) return BACS_OK;
  scale_multi_kernel<<<a.blk0[a.cnt], 256, 0, (cudaStream_t)stream>>>(a, g_dev);
  BACS_CHECK_LAUNCH("bacs_scale_inplace_multi");
  return BACS_OK;
}

int bacs_pack_state(const double* sums, const double* counts, int T, int D, const int64_t* confmat, int K,
                    double* packed, bacs_stream_t stream) {
  BACS_REQUIRE(packed && (T == 0 || (sums && counts)) && (K == 0 || confmat), "bacs_pack_state: null pointer");
  const int n = T * D + T + K * K;
  if (n == 0) return BACS_OK;
  pack_state_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sums, counts, T * D, T, confmat, K * K, packed);
  BACS_CHECK_LAUNCH("bacs_pack_state");
  return BACS_OK;
}

int bacs_unpack_state(const double* packed, int T, int D, double* sums, double* counts, int64_t* confmat, int K,
                      bacs_stream_t stream) {
  BACS_REQUIRE(packed && (T == 0 || (sums && counts)) && (K == 0 || confmat), "bacs_unpack_state: null pointer");
  const int n = T * D + T + K * K;
  if (n == 0) return BACS_OK;
  unpack_state_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(packed, T * D, T, sums, counts, confmat,
                                                                        K * K);
  BACS_CHECK_LAUNCH("bacs_unpack_state");
  return BACS_OK;
}

/* out[0] = sum_i coef[i] * src[i][idx[i]] / (den[i] ? den[i][didx[i]] : 1)   (n <= 8 terms)
 * -- assembles the step's loss scalar on the device from the fp64 accumulators. */
int bacs_combine_scalars(int n, const double* const* src, const int* idx, const double* const* den, const int* didx,
                         const float* coef, float* out, bacs_stream_t stream) {
  BACS_REQUIRE(n > 0 && n <= 8 && src && idx && coef && out, "bacs_combine_scalars: bad arguments");
  CombineArgs a;
  a.n = n;
  for (int i = 0; i < n; ++i) {
    a.t[i].src = src[i];
    a.t[i].idx = idx[i];
    a.t[i].den = den ? den[i] : nullptr;
    a.t[i].didx = didx ? didx[i] : 0;
    a.t[i].coef = coef[i];
    BACS_REQUIRE(src[i], "bacs_combine_scalars: null source %d", i);
  }
  combine_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a, out);
  BACS_CHECK_LAUNCH("bacs_combine_scalars");
  return BACS_OK;
}

}  // extern "C"
