// Cross-GPU combination of the per-step state over NVLink peer memory, fused with the prototype update.
//
// The only cross-rank state of the BACS loss path is a few tens of KB per step (per-task feature sums and
// pixel counts; the confusion matrix at evaluation): latency-bound, so instead of a ring / tree collective every
// rank reads all peers' buffers directly ("one-shot" all-reduce) and applies the running-mean update
// (loss/prototypes.py:158-163) in the same launch:
//   1. copy my packed fp64 state into my slot of the symmetric buffer (double-buffered by step parity),
//   2. publish: fence.sys + release-store of the step number into every peer's flag row,
//   3. wait until every peer's flag for me shows this step (acquire loads, bounded spin),
//   4. sum the W peer buffers in RANK ORDER (bit-identical result on every rank), write it back to `packed`,
//   5. (optional) proto[g] = (S[g] + cnt[g] proto[g]) / (cnt[g] + N[g]), cnt[g] += N[g]; ready flag.
// A peer can only be one step ahead (it needs my flag of step e+1 to finish step e+1), and it then writes the
// OTHER parity slot, so reads of step e never race with writes of step e+1.
#include "common.cuh"

namespace bacs {

constexpr int kPeerMaxWorld = 16;
struct PeerPtrs {
  double* buf[kPeerMaxWorld];         // peer r's symmetric region: [2][n_max] doubles
  unsigned int* flag[kPeerMaxWorld];  // peer r's flag row: [kPeerMaxWorld] step numbers, one per source rank
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024) peer_allreduce_update_kernel(PeerPtrs pp, int rank, int world, int n, int n_max,
                                                                     double* __restrict__ packed,
                                                                     unsigned int* __restrict__ step_dev,
                                                                     int* __restrict__ error_dev,
                                                                     // optional fused prototype update
                                                                     float* __restrict__ proto, void* __restrict__ count,
                                                                     int count_is_int64, int Tn, int D,
                                                                     int32_t* __restrict__ ready) {
  __shared__ int s_timeout;
  const unsigned int step = *step_dev + 1u;
  const int slot = (int)(step & 1u);
  double* mine = pp.buf[rank] + (size_t)slot * n_max;
  if (threadIdx.x == 0) s_timeout = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = packed[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    st_release_sys(pp.flag[threadIdx.x] + rank, step);  // tell peer threadIdx.x that my slot is filled
    // wait for that peer's slot (bounded: a missing peer must not hang the GPU)
    const unsigned int* f = pp.flag[rank] + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - step) < 0) {
      if (clock64() - t0 > 4000000000LL) {  // ~2 s
        s_timeout = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_timeout) {
    if (threadIdx.x == 0 && error_dev) *error_dev = 1;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(pp.buf[r] + (size_t)slot * n_max + i);
    packed[i] = s;
  }
  if (threadIdx.x == 0) *step_dev = step;
  if (proto == nullptr) return;
  __syncthreads();  // packed[] (global, this block) is complete
  // ---- fused running-mean update (same arithmetic as proto_update_kernel) ----
  __shared__ float s_old[64], s_den[64];
  __shared__ int s_upd[64];
  __shared__ int s_nonzero;
  const double* sums = packed;
  const double* counts = packed + (size_t)Tn * D;
  if (threadIdx.x == 0) s_nonzero = 0;
  __syncthreads();
  if (threadIdx.x < Tn) {
    const int g = threadIdx.x;
    const double nn_d = counts[g];
    float oldc, den;
    int nz;
    if (count_is_int64) {
      int64_t* c = reinterpret_cast<int64_t*>(count);
      const int64_t o = c[g];
      const int64_t nn = o + (int64_t)nn_d;
      oldc = (float)o;
      den = (float)nn;
      if (nn_d > 0) c[g] = nn;
      nz = (nn_d > 0 ? nn : o) != 0;
    } else {
      float* c = reinterpret_cast<float*>(count);
      const float o = c[g];
      const float nn = __fadd_rn(o, (float)nn_d);
      oldc = o;
      den = nn;
      if (nn_d > 0) c[g] = nn;
      nz = (nn_d > 0 ? nn : o) != 0.f;
    }
    s_old[g] = oldc;
    s_den[g] = den;
    s_upd[g] = nn_d > 0;
    if (nz) atomicAdd(&s_nonzero, 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tn * D; i += blockDim.x) {
    const int g = i / D;
    if (s_upd[g]) {
      const float num = __fadd_rn((float)sums[i], __fmul_rn(s_old[g], proto[i]));
      proto[i] = __fdiv_rn(num, s_den[g]);
    }
  }
  if (threadIdx.x == 0 && ready) *ready = (s_nonzero == Tn) ? 1 : 0;
}

}  // namespace bacs

using namespace bacs;

extern "C" {

int bacs_peer_allreduce(double* packed, int n, int n_max, int rank, int world, const uint64_t* peer_buf_host,
                        const uint64_t* peer_flag_host, uint32_t* step_dev, int32_t* error_dev, float* proto,
                        void* count, int count_is_int64, int T, int D, int32_t* ready, bacs_stream_t stream) {
  BACS_REQUIRE(packed && peer_buf_host && peer_flag_host && step_dev, "bacs_peer_allreduce: null pointer");
  BACS_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bacs_peer_allreduce: bad rank/world");
  BACS_REQUIRE(n > 0 && n <= n_max, "bacs_peer_allreduce: n=%d exceeds the symmetric buffer (%d)", n, n_max);
  if (proto) BACS_REQUIRE(count && T > 0 && T <= 64 && D > 0 && n >= T * D + T, "bacs_peer_allreduce: bad prototype shape");
  PeerPtrs pp;
  for (int r = 0; r < kPeerMaxWorld; ++r) {
    pp.buf[r] = r < world ? reinterpret_cast<double*>(peer_buf_host[r]) : nullptr;
    pp.flag[r] = r < world ? reinterpret_cast<unsigned int*>(peer_flag_host[r]) : nullptr;
  }
  peer_allreduce_update_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pp, rank, world, n, n_max, packed, step_dev, error_dev,
                                                                    proto, count, count_is_int64, T, D, ready);
  BACS_CHECK_LAUNCH("bacs_peer_allreduce");
  return BACS_OK;
}

}  // extern "C"
