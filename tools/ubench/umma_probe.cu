// Probe of the tcgen05 plumbing the tensor-core teacher-distill kernel relies on (sm_100a):
//   1. kind::tf32 MMA, A from TMEM (lane = row, column = k), B K-major / no swizzle from shared memory
//   2. the same shared-memory bytes read as an MN-major B operand (roles of N and K swapped)
//   3. cycles per small-N MMA, 4. tcgen05.st / tcgen05.ld throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int N, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
               "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(addr) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
#define TC_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TC_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define TC_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define TC_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

constexpr int KTOT = 112;           // kappa range of the B array (28 groups of 4)
constexpr int NROW = 32;            // rows: 0..15 "hi", 16..31 "lo"
constexpr int KF = 56;              // forward test contraction length (7 k-steps)
// byte offset of element (kappa, row) in the shared-memory B array
__host__ __device__ inline int boff(int kappa, int row) { return (kappa / 4) * 512 + (row / 8) * 128 + (row % 8) * 16 + (kappa % 4) * 4; }

// out1 [128][32]: forward-like   D = A[128 x 56] * B^T (rows 0..31 as N, kappa as K)
// out2 [128][112]: backward-like D = R[128 x 16] * B (kappa as N, rows 0..15 as K), B read MN-major
// cyc[0..]: timings
__global__ void __launch_bounds__(256, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ R, const float* __restrict__ Bsm,
                                                       float* __restrict__ out1, float* __restrict__ out2, long long* __restrict__ cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* bs = reinterpret_cast<float*>(smem);
  for (int k = tid; k < KTOT * NROW; k += blockDim.x) bs[k] = Bsm[k];   // already in the device layout
  if (tid == 0) {
    for (int k = 0; k < 4; ++k) mbar_init(&bar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA reads)
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int m = (warp & 3) * 32 + (tid & 31);   // TMEM lane = matrix row of this thread
  const uint32_t sb = smem_u32(smem);

  // ---------------- test 1: forward-like ----------------
  if (warp < 4) {
    uint32_t v[8];
    for (int c = 0; c < KF; c += 8) {
      for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(A[m * KF + c + e]);
      tmem_st8(tb + lane_base + c, v);
    }
    TC_WAIT_ST();
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  if (tid == 0) {
    for (int s = 0; s < KF / 8; ++s)
      umma_tf32_ts(tb + 128, tb + 8 * s, make_desc(sb + (2 * s) * 512, 512, 128), make_idesc(32, 0), s > 0);
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  TC_FENCE_AFTER();
  if (warp < 4) {
    uint32_t v[32];
    tmem_ld32(tb + lane_base + 128, v);
    TC_WAIT_LD();
    for (int e = 0; e < 32; ++e) out1[m * 32 + e] = __uint_as_float(v[e]);
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();

  // ---------------- test 2: backward-like (MN-major B) ----------------
  if (warp < 4) {
    uint32_t v[8];
    for (int c = 0; c < 16; c += 8) {
      for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(R[m * 16 + c + e]);
      tmem_st8(tb + lane_base + c, v);
    }
    TC_WAIT_ST();
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  if (tid == 0) {
    for (int t = 0; t < 2; ++t)
      umma_tf32_ts(tb + 256, tb + 8 * t, make_desc(sb + t * 128, /*LBO (K groups)*/ 128, /*SBO (MN groups)*/ 512), make_idesc(112, 1), t > 0);
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  TC_FENCE_AFTER();
  if (warp < 4) {
    uint32_t v[32];
    for (int c = 0; c < 112; c += 32) {
      tmem_ld32(tb + lane_base + 256 + c, v);
      TC_WAIT_LD();
      for (int e = 0; e < 32 && c + e < 112; ++e) out2[m * 112 + c + e] = __uint_as_float(v[e]);
    }
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();

  // ---------------- test 3: MMA issue rates ----------------
  const int reps = 512;
  if (tid == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int s = r % 7;
      umma_tf32_ts(tb + 128, tb + 8 * s, make_desc(sb + (2 * s) * 512, 512, 128), make_idesc(32, 0), 1);
      umma_tf32_ts(tb + 128, tb + 56 + 8 * s, make_desc(sb + (2 * s) * 512, 512, 128), make_idesc(16, 0), 1);
    }
    umma_commit(&bar[2]);
    long long t1 = clock64();
    mbar_wait(&bar[2], 0);
    long long t2 = clock64();
    cyc[0] = t1 - t0;   // issue time of 2*reps MMAs (N=32 + N=16)
    cyc[1] = t2 - t0;   // until completion
    TC_FENCE_AFTER();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int t = r & 1;
      umma_tf32_ts(tb + 256, tb + 8 * t, make_desc(sb + t * 128, 128, 512), make_idesc(112, 1), 1);
    }
    umma_commit(&bar[3]);
    t1 = clock64();
    mbar_wait(&bar[3], 0);
    t2 = clock64();
    cyc[2] = t1 - t0;
    cyc[3] = t2 - t0;   // reps MMAs of N=112 (MN-major B)
    TC_FENCE_AFTER();
  }
  __syncthreads();
  TC_FENCE_AFTER();

  // ---------------- test 4: tcgen05.st / ld throughput, 4 and 8 warps ----------------
  for (int nwarps = 4; nwarps <= 8; nwarps += 4) {
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
      uint32_t v[32];
      for (int e = 0; e < 32; ++e) v[e] = tid + e;
      const uint32_t base = tb + lane_base + (warp >> 2) * 256;
      for (int r = 0; r < 64; ++r) tmem_st32(base + (r & 7) * 32, v);
      TC_WAIT_ST();
    }
    __syncthreads();
    long long t1 = clock64();
    uint32_t acc = 0;
    if (warp < nwarps) {
      uint32_t v[32];
      const uint32_t base = tb + lane_base + (warp >> 2) * 256;
      for (int r = 0; r < 64; ++r) {
        tmem_ld32(base + (r & 7) * 32, v);
        TC_WAIT_LD();
        for (int e = 0; e < 32; ++e) acc += v[e];
      }
    }
    __syncthreads();
    long long t2 = clock64();
    if (tid == 0) {
      cyc[4 + (nwarps / 4 - 1) * 2] = t1 - t0;   // 64 x (nwarps x 32 lanes x 32 cols x 4 B) stored
      cyc[5 + (nwarps / 4 - 1) * 2] = t2 - t1;
    }
    if (acc == 0x12345678u) out1[0] = 1.f;
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}


// ---- layout decoder: A = identity on the first 16 rows, B element value encodes its (kappa,row) coordinates ----
__global__ void __launch_bounds__(128, 1) decode_kernel(const float* __restrict__ Bsm, float* __restrict__ out, int lbo, int sbo, int N, int bmn,
                                                        int start_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[1];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* bs = reinterpret_cast<float*>(smem);
  for (int k = tid; k < KTOT * NROW; k += blockDim.x) bs[k] = Bsm[k];
  if (tid == 0) { mbar_init(&bar[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  TC_FENCE_BEFORE(); __syncthreads(); TC_FENCE_AFTER();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int m = tid;
  uint32_t v[8];
  for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(m == e ? 1.f : 0.f);   // K = 8: row m < 8 picks k = m
  tmem_st8(tb + lane_base, v);
  TC_WAIT_ST();
  TC_FENCE_BEFORE(); __syncthreads(); TC_FENCE_AFTER();
  if (tid == 0) {
    umma_tf32_ts(tb + 256, tb, make_desc(smem_u32(smem) + start_off, lbo, sbo), make_idesc(N, bmn), 0);
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  TC_FENCE_AFTER();
  uint32_t r[32];
  for (int c = 0; c < N; c += 32) {
    tmem_ld32(tb + lane_base + 256 + c, r);
    TC_WAIT_LD();
    for (int e = 0; e < 32 && c + e < N; ++e) out[m * 256 + c + e] = __uint_as_float(r[e]);
  }
  TC_FENCE_BEFORE(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

// ---- MMA rate: 16 unrolled MMAs with precomputed descriptors, repeated ----
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* __restrict__ cyc, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[1];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* bs = reinterpret_cast<float*>(smem);
  for (int k = tid; k < 16384; k += blockDim.x) bs[k] = 0.f;
  if (tid == 0) { mbar_init(&bar[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  TC_FENCE_BEFORE(); __syncthreads(); TC_FENCE_AFTER();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    const uint64_t d0 = make_desc(smem_u32(smem), 512, 128);
    const uint32_t id = make_idesc(N, 0);
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (leader) {
#pragma unroll
        for (int s = 0; s < 16; ++s) umma_tf32_ts(tb + 256, tb + 8 * (s & 7), d0 + (uint64_t)((s & 7) * 64), id, 1);
      }
      __syncwarp();
    }
    if (leader) umma_commit(&bar[0]);
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    if (leader) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  }
  TC_FENCE_BEFORE(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

// ---- issue cost of the forward pattern of distill_tc: 7 x (N=32, N=16) + commit per block ----
__global__ void __launch_bounds__(128, 1) pattern_kernel(long long* __restrict__ cyc, int reps, int with_commit, int wait_each) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* bs = reinterpret_cast<float*>(smem);
  for (int k = tid; k < 16384; k += blockDim.x) bs[k] = 0.f;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  TC_FENCE_BEFORE(); __syncthreads(); TC_FENCE_AFTER();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t sb = smem_u32(smem);
    long long issue = 0;
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      const long long ta = clock64();
      if (leader) {
        const uint32_t tw = tb + (r & 1) * 256, va = tw + ((r >> 1) & 1) * 112;
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          const uint64_t bd = make_desc(sb + s * 1024, 512, 128);
          umma_tf32_ts(tw + 224, va + 8 * s, bd, make_idesc(32, 0), (r | s) != 0);
          umma_tf32_ts(tw + 224 + 16, va + 56 + 8 * s, bd, make_idesc(16, 0), 1);
        }
        if (with_commit) umma_commit(&bar[0]);
      }
      __syncwarp();
      issue += clock64() - ta;
      if (with_commit && wait_each) { mbar_wait(&bar[0], ph); ph ^= 1; }
    }
    if (leader) umma_commit(&bar[1]);
    __syncwarp();
    mbar_wait(&bar[1], 0);
    const long long t2 = clock64();
    if (leader) { cyc[0] = issue; cyc[1] = t2 - t0; }
  }
  TC_FENCE_BEFORE(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  srand(1);
  auto rnd = [] { return tf32r((float)rand() / RAND_MAX * 2.f - 1.f); };
  std::vector<float> A(128 * KF), R(128 * 16), Blog(KTOT * NROW), Bdev(KTOT * NROW);
  for (auto& x : A) x = rnd();
  for (auto& x : R) x = rnd();
  for (int k = 0; k < KTOT; ++k)
    for (int r = 0; r < NROW; ++r) {
      const float v = rnd();
      Blog[k * NROW + r] = v;
      Bdev[boff(k, r) / 4] = v;
    }
  float *dA, *dR, *dB, *o1, *o2;
  long long* dc;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dR, R.size() * 4)); CK(cudaMalloc(&dB, Bdev.size() * 4));
  CK(cudaMalloc(&o1, 128 * 32 * 4)); CK(cudaMalloc(&o2, 128 * 112 * 4)); CK(cudaMalloc(&dc, 16 * 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, R.data(), R.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bdev.data(), Bdev.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(o1, 0, 128 * 32 * 4)); CK(cudaMemset(o2, 0, 128 * 112 * 4)); CK(cudaMemset(dc, 0, 128));
  const int smem = KTOT * NROW * 4;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 256, smem>>>(dA, dR, dB, o1, o2, dc);
  CK(cudaDeviceSynchronize());
  std::vector<float> h1(128 * 32), h2(128 * 112);
  long long hc[16];
  CK(cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc, dc, 128, cudaMemcpyDeviceToHost));
  double e1 = 0, e2 = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      double s = 0;
      for (int k = 0; k < KF; ++k) s += (double)A[m * KF + k] * Blog[k * NROW + n];
      e1 = fmax(e1, fabs(s - h1[m * 32 + n]));
    }
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 112; ++k) {
      double s = 0;
      for (int y = 0; y < 16; ++y) s += (double)R[m * 16 + y] * Blog[k * NROW + y];
      e2 = fmax(e2, fabs(s - h2[m * 112 + k]));
    }
  printf("test1 fwd-like  (A in TMEM, B K-major)  max abs err %.3e  (sample %f)\n", e1, h1[5]);
  printf("test2 bwd-like  (A in TMEM, B MN-major) max abs err %.3e  (sample %f)\n", e2, h2[7]);
  printf("test3 1024 MMAs (N=32,N=16 pairs): issue %lld cyc, done %lld cyc -> %.1f cyc per pair\n", hc[0], hc[1], hc[1] / 512.0);
  printf("test3 512 MMAs N=112 MN-major: issue %lld cyc, done %lld cyc -> %.1f cyc each\n", hc[2], hc[3], hc[3] / 512.0);
  for (int i = 0; i < 2; ++i) {
    const double bytes = 64.0 * (4 * (i + 1)) * 32 * 32 * 4;
    printf("test4 %d warps: st %lld cyc (%.1f B/cyc), ld+wait %lld cyc (%.1f B/cyc)\n", 4 * (i + 1), hc[4 + 2 * i], bytes / hc[4 + 2 * i], hc[5 + 2 * i], bytes / hc[5 + 2 * i]);
  }

  // ---- decode the MN-major read pattern ----
  {
    std::vector<float> Bc(KTOT * NROW, 0.f);
    for (int k = 0; k < KTOT; ++k)
      for (int r = 0; r < 16; ++r) Bc[boff(k, r) / 4] = (float)(k * 16 + r);
    float *dBc, *dout;
    CK(cudaMalloc(&dBc, Bc.size() * 4)); CK(cudaMalloc(&dout, 128 * 256 * 4));
    CK(cudaMemcpy(dBc, Bc.data(), Bc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    struct V { int lbo, sbo, N, bmn, off; const char* name; };
    V vars[] = {{512, 128, 32, 0, 0, "K-major ref"}, {128, 512, 112, 0, 0, "K-major N=112"}};
    for (auto& vv : vars) {
      CK(cudaMemset(dout, 0, 128 * 256 * 4));
      decode_kernel<<<1, 128, smem>>>(dBc, dout, vv.lbo, vv.sbo, vv.N, vv.bmn, vv.off);
      CK(cudaDeviceSynchronize());
      std::vector<float> ho(128 * 256);
      CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
      printf("decode %-18s N=%3d: D[k][n] -> (kappa,row) read for B(n,k):\n", vv.name, vv.N);
      for (int k = 0; k < 8; k += 1) {
        printf("  k=%d:", k);
        for (int n = 0; n < 20; ++n) { int c = (int)ho[k * 256 + n]; printf(" (%d,%d)", c / 16, c % 16); }
        printf(" ... n=%d: ", vv.N - 1); { int c = (int)ho[k * 256 + vv.N - 1]; printf("(%d,%d)\n", c / 16, c % 16); }
      }
    }
  }
  // ---- MMA rates ----
  {
    long long* dc2; CK(cudaMalloc(&dc2, 64));
    long long h[2];
#define RATE(NN) { CK(cudaFuncSetAttribute(rate_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); \
      rate_kernel<NN><<<1, 128, 65536>>>(dc2, 64); CK(cudaDeviceSynchronize()); rate_kernel<NN><<<1, 128, 65536>>>(dc2, 64); CK(cudaDeviceSynchronize()); \
      CK(cudaMemcpy(h, dc2, 16, cudaMemcpyDeviceToHost)); \
      printf("rate N=%3d: 1024 MMAs issue %lld done %lld cyc -> %.1f cyc/MMA\n", NN, h[0], h[1], h[1] / 1024.0); }
    RATE(16) RATE(32) RATE(48) RATE(64) RATE(112) RATE(128) RATE(256)
  }

  {
    long long* dc3; CK(cudaMalloc(&dc3, 64));
    long long h[2];
    CK(cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int mode = 0; mode < 3; ++mode) {
      const int wc = mode > 0, we = mode > 1;
      pattern_kernel<<<1, 128, 65536>>>(dc3, 256, wc, we); CK(cudaDeviceSynchronize());
      pattern_kernel<<<1, 128, 65536>>>(dc3, 256, wc, we); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, dc3, 16, cudaMemcpyDeviceToHost));
      printf("pattern (7 x (N=32,N=16)%s%s): issue %.1f cyc per block, total %.1f cyc per block\n", wc ? " + commit" : "", we ? ", wait each" : "", h[0] / 256.0, h[1] / 256.0);
    }
  }
  printf("%s\n", (e1 < 1e-4 && e2 < 1e-4) ? "PROBE OK" : "PROBE MISMATCH");
  return 0;
}
