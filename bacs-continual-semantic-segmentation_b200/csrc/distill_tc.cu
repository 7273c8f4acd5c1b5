// Teacher distillation (loss/bacs_loss.py:258-294) on the 5th-generation tensor cores: GEMM form, tcgen05 + TMEM.
//
//   L = coef * sum_{b,a,Y} sqrt(S[b,a,Y]),  S = sum_X m(Y,X) (U(old)^2 - U(new)^2)^2,  U = bilinear x16 (align_corners=False)
//
// Inside one low-res cell (source rows i,i+1; columns j,j+1) both up-sampled maps are bilinear in (ty,tx), so with
// p = old - new, q = old + new at the four corners
//   U(old)^2 - U(new)^2 = P*Q = sum_{r,s<3} E_rs b2_r(ty) b2_s(tx),      E = conv2(p, q)            (3x3)
//   (P*Q)^2             = sum_{q,k<5} V_qk b4_q(ty) b4_k(tx),            V = conv2(E, E)            (5x5)
// in the un-normalised Bernstein bases b^d_r(t) = (1-t)^(d-r) t^r (no binomial factors appear in products; every
// basis function is >= 0 on [0,1], which keeps the expansion well conditioned).  Hence
//   S[ch, Y] = sum_{cell j} sum_{q,k} V_qk(ch, j) * [ b4_q(ty_Y) * M_k(Y, j) ],   M_k(Y,j) = sum_{X in cell j} m(Y,X) b4_k(tx_X)
// is a GEMM: S[128 channels x 16 rows] = V[128 x 28 w] * Bm[28 w x 16] per (image, source-row interval, 16-row block):
// V (per channel, 25 values per cell) is built by the CUDA cores straight into TENSOR MEMORY as the A operand, Bm
// (channel independent, from the mask) is built into shared memory, and the contraction -- which the FMA kernel of
// distill.cu evaluates row by row, 16 times per cell -- runs on tcgen05.mma.  The backward is the transposed GEMM
//   W[ch, (j,q,k)] = sum_Y rs[ch,Y] * Bm[(j,q,k), Y],   rs = 1/sqrt(S)
// followed by the chain rule  dE_a = sum_b W_{a+b} E_b,  d new_cd = -2 sum dE_{c+c',d+d'} new_c'd'  on the CUDA cores.
//
// Precision: both operands are split x = hi + lo with hi = x & 0xffffe000 (exact in tf32) and three products
// (hi*hi, hi*lo, lo*hi) accumulate in fp32: the dropped lo*lo term and the tf32 truncation of lo are <= 2^-20
// relative per product; measured 1e-7 (loss) / 3e-7 of max|g| (gradient) against fp64, also for near-identical
// maps (tools/distill_tc_proto.py).  old == new gives p = 0, E = V = 0 and S = 0, gradient 0 exactly.
//
// Work decomposition: persistent CTAs, one per SM; a CTA walks a contiguous range of (image, 256-channel block,
// source-row interval) items top to bottom, so the gradient an interval sends to its lower source row is carried in
// shared memory; the first row of a range is completed by distill_tc_finish_kernel (a + b, order independent).
//   warps 0-3 / 4-7 : builder warpgroups, one thread per channel (TMEM lane = channel): V chunks -> TMEM, epilogue
//                     (row norms, rs -> TMEM), chain rule of the backward, gradient rows
//   warps 8-10      : mask moments and the Bm operand into a shared-memory ring
//   warp 11         : one elected lane issues every tcgen05.mma / commit and the TMA loads of the attention rows
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace bacs {
namespace dtc {

constexpr int kThreads = 384;
constexpr int kChanCta = 256;      // channels per item: 2 warpgroups x 128 TMEM lanes
constexpr int kRows = 16;          // rows of one block = N of the forward MMA, K of the backward MMA
constexpr int kPairK = 56;         // contraction slots of a pair of cells: 2 x 28 (25 used per cell), interleaved
constexpr int kSlotBytes = 14592;  // one ring slot: forward chunk (7168 B) or backward group (2 x 7296 B)
constexpr int kBwdLbo = 1824;      // backward layout: byte stride between groups of 4 rows (bank-spread, 16 B multiple)
constexpr int kBwdHalf = 4 * kBwdLbo;
constexpr int kCarryStride = 33;   // floats per thread-private carry row (conflict-free)
constexpr int kMaxW = 32;
constexpr int kColsWg = 256;       // TMEM columns of one warpgroup: 2 x 112 (V chunk / dV group) + 32 (S, then rs)
constexpr int kColVB = 112;
constexpr int kColDS = 224;
constexpr int kRing = 3;          // ring slots of the Bm operand: one per builder warp (a slot has ONE producer, so its
                                  // 'free' barrier is observed phase by phase)

struct Params {
  const uint8_t* mask;
  void* dnew;
  int B, A, h, w, H, W;
  int ncb, total_items;
  float sy, sx, grad_coef;
  double* partials;   // [grid] sum of row norms per CTA
  float* bnd_own;     // [grid][256][32] own-row partial of a range's first item (when it is not the top row)
  float* bnd_low;     // [grid][256][32] lower-row partial of a range's last item
  int* bnd_info;      // [grid][2] item range of every CTA
  float* mb_scratch;  // [grid][2][256][32] partial gradient rows of intervals with more than one row block
  int n_att;
  int want_grad;
  int box_chan;       // channels per TMA box: min(256, B*A)
  int knock;          // diagnostics only (BACS_DTC_KNOCK): 1 = operand builders skip their arithmetic, 2 = no MMAs are issued,
                      // 4 = builders skip the V / dE arithmetic; results are wrong, the timing shows the critical role
  int fast_mask;      // mask rows staged in shared memory by bulk copies + basis table (needs 16-byte aligned rows)
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
#ifdef BACS_DTC_DEBUG
// bounded waits: the first wait that does not complete records (CTA, thread, barrier offset, parity) and every later
// wait returns immediately, so a protocol bug ends the kernel with a trace instead of a hang
__device__ int g_dtc_dead = 0;
__device__ int g_dtc_trace[64 * 4];
__device__ int g_dtc_barbase = 0;
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; spin < (1ll << 22); ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok || *(volatile int*)&g_dtc_dead) return;
  }
  if ((threadIdx.x & 31) != 0) return;
  const int slot = atomicAdd(&g_dtc_dead, 1);
  if (slot < 63) {
    g_dtc_trace[4 * slot] = blockIdx.x;
    g_dtc_trace[4 * slot + 1] = threadIdx.x;
    g_dtc_trace[4 * slot + 2] = (int)smem_u32(b);
    g_dtc_trace[4 * slot + 3] = (int)parity;
  }
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(b)),
      "r"(parity)
      : "memory");
}
#endif
#ifdef BACS_DTC_PROFILE
__device__ long long g_dtc_prof[160 * 3 * 16];
#define PROF_DECL long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prof_t0 = clock64(); (void)prof_t0
#define PROF_WAIT(cat, bar, par) do { const long long t__ = clock64(); mbar_wait(bar, par); prof_acc[cat] += clock64() - t__; } while (0)
#define PROF_MARK(cat) do { const long long t__ = clock64(); prof_acc[cat] += t__ - prof_t0; prof_t0 = t__; } while (0)
#define PROF_DUMP(role) do { if ((threadIdx.x & 31) == 0) for (int k__ = 0; k__ < 16; ++k__) g_dtc_prof[(blockIdx.x * 3 + role) * 16 + k__] = prof_acc[k__]; } while (0)
#else
#define PROF_DECL
#define PROF_WAIT(cat, bar, par) mbar_wait(bar, par)
#define PROF_MARK(cat)
#define PROF_DUMP(role)
#endif
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
// D[tmem] (+)= A[tmem, 128 lanes x 8 columns of tf32] * B[smem descriptor, N x 8]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
// the same with the descriptor as two words: stepping the start address is one 32-bit add on the low word
__device__ __forceinline__ void umma_tf32_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc,
                                              uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 d;\nmov.b64 d, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d, %4, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, no swizzle: element (n, k) at (n/8)*sbo + (n%8)*16 + (k/4)*lbo + (k%4)*4 bytes (tools/ubench/umma_probe.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int N) {  // tf32 x tf32 -> fp32, M = 128, both operands K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
#define TC_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TC_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define TC_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define TC_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

__device__ __forceinline__ void tmem_st8(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t a, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t a, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(a),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(a)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t a, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(a)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t a, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(a)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ uint32_t lo32(F2 a) { return (uint32_t)a.v; }
__device__ __forceinline__ uint32_t hi32(F2 a) { return (uint32_t)(a.v >> 32); }
__device__ __forceinline__ F2 pack32(uint32_t lo, uint32_t hi) {
  F2 r;
  r.v = (unsigned long long)lo | ((unsigned long long)hi << 32);
  return r;
}
__device__ __forceinline__ F2 tf32_hi2(F2 a) {
  F2 r;
  r.v = a.v & 0xFFFFE000FFFFE000ull;
  return r;
}
__device__ __forceinline__ float rsqrt_fast_tc(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// un-normalised Bernstein basis of degree 4 at t: (1-t)^(4-k) t^k
__device__ __forceinline__ void bern4(float t, float* o) {
  const float s = 1.f - t, s2 = s * s, t2 = t * t, st = s * t;
  o[0] = s2 * s2;
  o[1] = s2 * st;
  o[2] = s2 * t2;
  o[3] = st * t2;
  o[4] = t2 * t2;
}

// ---------------------------------------------------------------- shared-memory plan (host + device)
// moments of a unit are re-read by its backward jobs; narrow maps need two buffers (jobs of consecutive units overlap)
__host__ __device__ inline int mom_buffers(int w) { return w >= 24 ? 1 : 2; }
struct SmemPlan {
  int att_off, att_slot_bytes;   // n_att slots x {old rows, new rows}: 256 rows of row_bytes each
  int ring_off, carry_off, mom_off, yw_off, xw_off, rowstart_off, colstart_off, blkpfx_off, bar_off, phi_off, mrow_off, total;
};
__host__ __device__ inline SmemPlan make_plan(int n_att, int fast_mask, int row_bytes, int H, int W, int h, int w) {
  SmemPlan s;
  int o = 0;
  s.att_off = o;
  s.att_slot_bytes = 2 * kChanCta * row_bytes;
  o += n_att * s.att_slot_bytes;
  s.ring_off = o;
  o += kRing * kSlotBytes;
  o = (o + 15) & ~15;
  s.carry_off = o;
  o += kChanCta * kCarryStride * 4;
  s.mom_off = o;
  o += mom_buffers(w) * kRows * w * 5 * 4;
  s.yw_off = o;
  o += H * 4;
  s.xw_off = o;
  o += W * 4;
  s.rowstart_off = o;
  o += (h + 1) * 4;
  s.colstart_off = o;
  o += (w + 1) * 4;
  s.blkpfx_off = o;
  o += (h + 1) * 4;
  o = (o + 15) & ~15;
  s.bar_off = o;
  o += 64 * 8;
  s.phi_off = o;
  if (fast_mask) o += 5 * W * 4;
  s.mrow_off = o;
  if (fast_mask) o += 2 * kRows * (W + 16);
  s.total = o;
  return s;
}

// barrier indices
enum {
  BAR_VFULL = 0,   // [wg][buf]  V chunk stored to TMEM (128 arrivals)
  BAR_VFREE = 4,   // [wg][buf]  MMAs that read the chunk are done (commit)
  BAR_WREADY = 8,  // [wg][buf]  dV group complete in TMEM (commit)
  BAR_WDONE = 12,  // [wg][buf]  dV group read (128 arrivals)
  BAR_SFULL = 16,  // [wg]       S accumulators complete (commit)
  BAR_RFULL = 18,  // [wg]       rs stored to TMEM (128 arrivals)
  BAR_BFULL = 20,  // [slot]     Bm chunk / group built (1 arrival)
  BAR_BFREE = 28,  // [slot]     MMAs that read the slot are done (commit)
  BAR_ATT = 36,    // [slot]     attention rows landed (TMA tx)
  BAR_ITEM = 40,   //            an item is finished by all 256 builder threads
  BAR_MASK = 41,   // [2]        mask rows of a unit landed (bulk copy tx)
  BAR_COUNT = 43
};

// attention-row slots: which slot holds the upper / lower source row of an item (same state machine on both sides)
struct AttPlan {
  int slotU, slotW;
  int newU, newW;  // slot that needs a fresh load for this item, or -1
};
__device__ __forceinline__ AttPlan att_next(const AttPlan& prev, bool first, int i, int h, int n_att) {
  AttPlan a;
  const bool clampW = (i + 1 > h - 1);  // lower row == upper row
  if (first) {
    a.slotU = 0;
    a.newU = 0;
    a.slotW = clampW ? 0 : 1;
    a.newW = clampW ? -1 : 1;
    return a;
  }
  if (i > 0) {  // same segment: the previous lower row is the upper row now
    a.slotU = prev.slotW;
    a.newU = -1;
    if (clampW) {
      a.slotW = a.slotU;
      a.newW = -1;
    } else {
      a.slotW = (n_att == 3) ? (3 - prev.slotU - prev.slotW) : prev.slotU;
      a.newW = a.slotW;
    }
    return a;
  }
  // new segment: the previous item was the last interval of its segment and holds one slot only
  const int used = prev.slotU;
  a.slotU = (used + 1) % n_att;
  a.newU = a.slotU;
  if (clampW) {
    a.slotW = a.slotU;
    a.newW = -1;
  } else {
    a.slotW = (n_att == 3) ? (used + 2) % 3 : used;
    a.newW = a.slotW;
  }
  return a;
}

// ---------------------------------------------------------------- compile-time polynomial products
// V_QK = sum over E_a E_b with a + b = (Q,K); E2 = 2 E exploits the symmetry (45 products instead of 81)
template <int Q, int K>
__device__ __forceinline__ F2 vcoef(const F2* E, const F2* E2) {
  F2 acc = f2b(0.f);
  bool first = true;
#pragma unroll
  for (int ia = 0; ia < 9; ++ia) {
#pragma unroll
    for (int ib = ia; ib < 9; ++ib) {
      if ((ia / 3) + (ib / 3) == Q && (ia % 3) + (ib % 3) == K) {
        const F2 x = (ia == ib) ? E[ia] : E2[ia];
        acc = first ? mul2(x, E[ib]) : fma2(x, E[ib], acc);
        first = false;
      }
    }
  }
  return acc;
}
// dE_a = sum_b W_{a+b} E_b
template <int R, int S>
__device__ __forceinline__ F2 decoef(const F2* Wv, const F2* E) {
  F2 acc = f2b(0.f);
  bool first = true;
#pragma unroll
  for (int ib = 0; ib < 9; ++ib) {
    const int e = (R + ib / 3) * 5 + (S + ib % 3);
    acc = first ? mul2(Wv[e], E[ib]) : fma2(Wv[e], E[ib], acc);
    first = false;
  }
  return acc;
}

template <typename T>
struct RowIO;
template <>
struct RowIO<float> {
  // 8 columns starting at j0 of row r (swizzled TMA layout), as floats
  __device__ static __forceinline__ void load8(const uint8_t* arr, int r, int j0, int rb, uint32_t swz, float* o) {
    const uint32_t off = (uint32_t)r * rb + (uint32_t)j0 * 4;
    const uint32_t o0 = off ^ (((off >> 7) & swz) << 4), o1 = (off + 16) ^ ((((off + 16) >> 7) & swz) << 4);
    const float4 a = *reinterpret_cast<const float4*>(arr + o0);
    const float4 b = *reinterpret_cast<const float4*>(arr + o1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  __device__ static __forceinline__ float load1(const uint8_t* arr, int r, int j, int rb, uint32_t swz) {
    const uint32_t off = (uint32_t)r * rb + (uint32_t)j * 4;
    return *reinterpret_cast<const float*>(arr + (off ^ (((off >> 7) & swz) << 4)));
  }
  __device__ static __forceinline__ void store8(float* dst, const float* v) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <typename T16>
struct RowIO16 {
  __device__ static __forceinline__ void load8(const uint8_t* arr, int r, int j0, int rb, uint32_t swz, float* o) {
    const uint32_t off = (uint32_t)r * rb + (uint32_t)j0 * 2;
    const uint4 a = *reinterpret_cast<const uint4*>(arr + (off ^ (((off >> 7) & swz) << 4)));
    const T16* p = reinterpret_cast<const T16*>(&a);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = DT<T16>::to_f(p[k]);
  }
  __device__ static __forceinline__ float load1(const uint8_t* arr, int r, int j, int rb, uint32_t swz) {
    const uint32_t off = (uint32_t)r * rb + (uint32_t)j * 2;
    return DT<T16>::to_f(*reinterpret_cast<const T16*>(arr + (off ^ (((off >> 7) & swz) << 4))));
  }
  __device__ static __forceinline__ void store8(T16* dst, const float* v) {
    uint4 a;
    T16* p = reinterpret_cast<T16*>(&a);
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = DT<T16>::from_f(v[k]);
    *reinterpret_cast<uint4*>(dst) = a;
  }
};
template <>
struct RowIO<__nv_bfloat16> : RowIO16<__nv_bfloat16> {};
template <>
struct RowIO<__half> : RowIO16<__half> {};


// Packed corner pairs of 8 consecutive cells straight from the swizzled row: L[m] = (c[2m], c[2m+1]), R[m] = (c[2m+1], c[2m+2]);
// the column after the last one of the map is the last column itself.
template <typename T>
struct PairIO {
  __device__ static __forceinline__ void load(const uint8_t* arr, int r, int j0, bool last, int rb, uint32_t swz, F2* L, F2* R) {
    float c[9];
    RowIO<T>::load8(arr, r, j0, rb, swz, c);
    c[8] = last ? c[7] : RowIO<T>::load1(arr, r, j0 + 8, rb, swz);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      L[m] = f2(c[2 * m], c[2 * m + 1]);
      R[m] = f2(c[2 * m + 1], c[2 * m + 2]);
    }
  }
};
template <>
struct PairIO<__nv_bfloat16> {  // bf16 -> fp32 is a shift: both pair sets come from the packed words with no moves
  __device__ static __forceinline__ void load(const uint8_t* arr, int r, int j0, bool last, int rb, uint32_t swz, F2* L, F2* R) {
    const uint32_t off = (uint32_t)r * rb + (uint32_t)j0 * 2;
    const uint4 a = *reinterpret_cast<const uint4*>(arr + (off ^ (((off >> 7) & swz) << 4)));
    const uint32_t wv[4] = {a.x, a.y, a.z, a.w};
    uint32_t nxt;
    if (last) {
      nxt = a.w >> 16;
    } else {
      const uint32_t o2 = off + 16;
      nxt = *reinterpret_cast<const uint16_t*>(arr + (o2 ^ (((o2 >> 7) & swz) << 4)));
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const uint32_t right = m < 3 ? wv[m < 3 ? m + 1 : 3] << 16 : nxt << 16;
      L[m] = pack32(wv[m] << 16, wv[m] & 0xffff0000u);
      R[m] = pack32(wv[m] & 0xffff0000u, right);
    }
  }
};

// ================================================================== the kernel
template <typename T>
__global__ void __launch_bounds__(kThreads, 1) distill_tc_kernel(const __grid_constant__ CUtensorMap map_old,
                                                                 const __grid_constant__ CUtensorMap map_new, Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzled TMA boxes: 1 KB aligned
  __shared__ uint32_t s_tmem_base;
  __shared__ double s_red[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = P.h, w = P.w, H = P.H, W = P.W;
  const int row_bytes = w * (int)sizeof(T);
  const uint32_t swz = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
  const SmemPlan sp = make_plan(P.n_att, P.fast_mask, row_bytes, H, W, h, w);
  float* s_carry = reinterpret_cast<float*>(smem + sp.carry_off);
  float* s_mom = reinterpret_cast<float*>(smem + sp.mom_off);
  float* s_yw = reinterpret_cast<float*>(smem + sp.yw_off);
  float* s_xw = reinterpret_cast<float*>(smem + sp.xw_off);
  int* s_rowstart = reinterpret_cast<int*>(smem + sp.rowstart_off);
  int* s_colstart = reinterpret_cast<int*>(smem + sp.colstart_off);
  int* s_blkpfx = reinterpret_cast<int*>(smem + sp.blkpfx_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bar_off);

  // ---- prologue: interpolation tables, barriers, tensor memory ----
  for (int Y = tid; Y < H; Y += kThreads) {
    const Lerp l = lerp_half_pixel(Y, h, P.sy);
    s_yw[Y] = (l.i1 == l.i0) ? 0.f : l.w1;
    if (Y == 0 || lerp_half_pixel(Y - 1, h, P.sy).i0 != l.i0) s_rowstart[l.i0] = Y;
  }
  for (int X = tid; X < W; X += kThreads) {
    const Lerp l = lerp_half_pixel(X, w, P.sx);
    s_xw[X] = (l.i1 == l.i0) ? 0.f : l.w1;
    if (X == 0 || lerp_half_pixel(X - 1, w, P.sx).i0 != l.i0) s_colstart[l.i0] = X;
  }
  if (P.fast_mask) {
    float* phi = reinterpret_cast<float*>(smem + sp.phi_off);
    for (int X = tid; X < W; X += kThreads) {
      const Lerp l = lerp_half_pixel(X, w, P.sx);
      float ph[5];
      bern4((l.i1 == l.i0) ? 0.f : l.w1, ph);
#pragma unroll
      for (int k = 0; k < 5; ++k) phi[k * W + X] = ph[k];
    }
  }
  if (tid == 0) {
    s_rowstart[h] = H;
    s_colstart[w] = W;
    mbar_init(&bars[BAR_MASK], 1);
    mbar_init(&bars[BAR_MASK + 1], 1);
    for (int k = 0; k < 2; ++k)
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars[BAR_VFULL + 2 * k + b], 128);
        mbar_init(&bars[BAR_VFREE + 2 * k + b], 1);
        mbar_init(&bars[BAR_WREADY + 2 * k + b], 1);
        mbar_init(&bars[BAR_WDONE + 2 * k + b], 128);
      }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&bars[BAR_SFULL + k], 1);
      mbar_init(&bars[BAR_RFULL + k], 128);
    }
    for (int k = 0; k < kRing; ++k) {
      mbar_init(&bars[BAR_BFULL + k], 1);
      mbar_init(&bars[BAR_BFREE + k], 1);
    }
    for (int k = 0; k < 4; ++k) mbar_init(&bars[BAR_ATT + k], 1);
    mbar_init(&bars[BAR_ITEM], 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 11) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  if (tid == 0) {  // prefix of row blocks per interval (weights of the work partition)
    int acc = 0;
    for (int i = 0; i < h; ++i) {
      s_blkpfx[i] = acc;
      acc += (s_rowstart[i + 1] - s_rowstart[i] + kRows - 1) / kRows;
    }
    s_blkpfx[h] = acc;
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  const uint32_t tmem = s_tmem_base;

  // ---- this CTA's contiguous range of items, balanced by row blocks, never splitting an interval ----
  const int blk_seg = s_blkpfx[h];
  const int64_t wtot = (int64_t)(P.total_items / h) * blk_seg;
  auto range_start = [&](int k) -> int {
    const int64_t t = (int64_t)k * wtot / gridDim.x;
    const int seg = (int)(t / blk_seg), r = (int)(t - (int64_t)seg * blk_seg);
    int i = 0;
    while (i < h && s_blkpfx[i] < r) ++i;
    return seg * h + i;
  };
  const int it0 = range_start(blockIdx.x), it1 = (blockIdx.x + 1 == gridDim.x) ? P.total_items : range_start(blockIdx.x + 1);
  pdl_wait();      // tables, barriers and tensor memory are set up while the kernel before this one drains
  pdl_trigger();
  if (tid == 0) {
    P.bnd_info[2 * blockIdx.x] = it0;
    P.bnd_info[2 * blockIdx.x + 1] = it1;
  }
#ifdef BACS_DTC_DEBUG
  if (tid == 0 && blockIdx.x == 0) g_dtc_barbase = (int)smem_u32(bars);
#endif
  const int NCH = w / 2, NG = w / 4;
  const bool want_grad = P.want_grad != 0;
  const int jobs_per_unit = NCH + (want_grad ? NG : 0);
  double loss_d = 0.0;

  if (warp < 8) {
    // =================================================== builder warpgroups
    const int wg = warp >> 2;
    const int lrow = wg * 128 + (warp & 3) * 32 + lane;  // row of this thread in the CTA's 256-channel block
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16) + wg * kColsWg;
    float* carry = s_carry + lrow * kCarryStride;
    float* scr = P.mb_scratch + ((size_t)blockIdx.x * 2 * kChanCta + lrow) * kMaxW;  // [0]: own, [+256*32]: low
    uint32_t n_unit = 0, n_att0 = 0, n_att1 = 0, n_att2 = 0;   // completed units; loads seen per attention slot
    float loss_f = 0.f;
    PROF_DECL;
    const float gscale = -2.f * P.grad_coef;
    AttPlan ap{0, 0, -1, -1};
    for (int it = it0; it < it1; ++it) {
      const int seg = it / h, i = it - seg * h, b = seg / P.ncb, cb = seg - b * P.ncb;
      const int ch = cb * kChanCta + lrow;
      const bool active = ch < P.A;
      ap = att_next(ap, it == it0, i, h, P.n_att);
      auto wait_att = [&](int slot) {
        uint32_t& n = slot == 0 ? n_att0 : (slot == 1 ? n_att1 : n_att2);
        PROF_WAIT(0, &bars[BAR_ATT + slot], n & 1);
        ++n;
      };
      if (ap.newU >= 0) wait_att(ap.newU);
      if (ap.newW >= 0) wait_att(ap.newW);
      const uint8_t* oU = smem + sp.att_off + ap.slotU * sp.att_slot_bytes;
      const uint8_t* nU = oU + kChanCta * row_bytes;
      const uint8_t* oW = smem + sp.att_off + ap.slotW * sp.att_slot_bytes;
      const uint8_t* nW = oW + kChanCta * row_bytes;
      const int Ybeg = s_rowstart[i], Yend = s_rowstart[i + 1];
      const int nblk = (Yend - Ybeg + kRows - 1) / kRows;
      const bool boundary_start = (it == it0) && (i > 0);  // the row above belongs to another CTA
      if (i == 0 || boundary_start) {
        for (int j = 0; j < w; ++j) carry[j] = 0.f;
      }
      // 8 cells + the right neighbour column of the four source rows as packed pairs: L[m] = (col 2m, col 2m+1) is the
      // left corner of the cell pair (2m, 2m+1), R[m] = (col 2m+1, col 2m+2) its right corner; p = old - new, q = old + new
      struct Corners { F2 p00[4], p01[4], p10[4], p11[4], q00[4], q01[4], q10[4], q11[4], n00[4], n01[4], n10[4], n11[4]; };
      auto load_corners = [&](int j0, Corners& c, bool want_n) {
        F2 oUL[4], oUR[4], nUL[4], nUR[4], oWL[4], oWR[4], nWL[4], nWR[4];
        const bool last = j0 + 8 >= w;
        PairIO<T>::load(oU, lrow, j0, last, row_bytes, swz, oUL, oUR);
        PairIO<T>::load(nU, lrow, j0, last, row_bytes, swz, nUL, nUR);
        PairIO<T>::load(oW, lrow, j0, last, row_bytes, swz, oWL, oWR);
        PairIO<T>::load(nW, lrow, j0, last, row_bytes, swz, nWL, nWR);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          c.p00[m] = sub2(oUL[m], nUL[m]); c.q00[m] = add2(oUL[m], nUL[m]);
          c.p01[m] = sub2(oUR[m], nUR[m]); c.q01[m] = add2(oUR[m], nUR[m]);
          c.p10[m] = sub2(oWL[m], nWL[m]); c.q10[m] = add2(oWL[m], nWL[m]);
          c.p11[m] = sub2(oWR[m], nWR[m]); c.q11[m] = add2(oWR[m], nWR[m]);
          if (want_n) { c.n00[m] = nUL[m]; c.n01[m] = nUR[m]; c.n10[m] = nWL[m]; c.n11[m] = nWR[m]; }
        }
      };
      auto make_E = [&](const Corners& c, int m, F2* E) {
        const F2 p00 = c.p00[m], p01 = c.p01[m], p10 = c.p10[m], p11 = c.p11[m];
        const F2 q00 = c.q00[m], q01 = c.q01[m], q10 = c.q10[m], q11 = c.q11[m];
        E[0] = mul2(p00, q00);
        E[1] = fma2(p00, q01, mul2(p01, q00));
        E[2] = mul2(p01, q01);
        E[3] = fma2(p00, q10, mul2(p10, q00));
        E[4] = fma2(p00, q11, fma2(p01, q10, fma2(p10, q01, mul2(p11, q00))));
        E[5] = fma2(p01, q11, mul2(p11, q01));
        E[6] = mul2(p10, q10);
        E[7] = fma2(p10, q11, mul2(p11, q10));
        E[8] = mul2(p11, q11);
      };

      for (int blk = 0; blk < nblk; ++blk) {
        // ------------------------------------------------ forward: V chunks -> tensor memory
        for (int j0 = 0; j0 < w; j0 += 8) {
          Corners cn;
          load_corners(j0, cn, false);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int cidx = (j0 >> 1) + m, buf = cidx & 1;
            F2 E[9], E2[9];
            make_E(cn, m, E);
#pragma unroll
            for (int k = 0; k < 9; ++k) E2[k] = add2(E[k], E[k]);
            const uint32_t vb = tm + buf * kColVB;
            uint32_t hi[32], lo[32];
            if (P.knock & 4) {
              const uint32_t nvc = n_unit * (uint32_t)(NCH >> 1) + (uint32_t)(cidx >> 1);
              if (nvc > 0) mbar_wait(&bars[BAR_VFREE + 2 * wg + buf], (nvc - 1) & 1);
              TC_FENCE_AFTER();
              for (int k = 0; k < 32; ++k) hi[k] = lo[k] = lo32(E[k % 9]);
              tmem_st32(vb, hi);
              TC_WAIT_ST();
              TC_FENCE_BEFORE();
              mbar_arrive(&bars[BAR_VFULL + 2 * wg + buf]);
              continue;
            }
#define BACS_V(e, Q, K)                         \
  {                                             \
    const F2 v = vcoef<Q, K>(E, E2);            \
    const F2 vh = tf32_hi2(v);                  \
    const F2 vl = sub2(v, vh);                  \
    hi[2 * ((e) & 15)] = lo32(vh);              \
    hi[2 * ((e) & 15) + 1] = hi32(vh);          \
    lo[2 * ((e) & 15)] = lo32(vl);              \
    lo[2 * ((e) & 15) + 1] = hi32(vl);          \
  }
            BACS_V(0, 0, 0) BACS_V(1, 0, 1) BACS_V(2, 0, 2) BACS_V(3, 0, 3) BACS_V(4, 0, 4)
            BACS_V(5, 1, 0) BACS_V(6, 1, 1) BACS_V(7, 1, 2) BACS_V(8, 1, 3) BACS_V(9, 1, 4)
            BACS_V(10, 2, 0) BACS_V(11, 2, 1) BACS_V(12, 2, 2) BACS_V(13, 2, 3) BACS_V(14, 2, 4)
            BACS_V(15, 3, 0)
            {
              // the buffer is free when the MMAs of the chunk two before this one are done; waiting HERE lets the 16
              // coefficients above overlap the tensor core's turnaround.  Chunk number on this buffer since the kernel
              // started: units * NCH/2 + cidx/2 (NCH is even)
              const uint32_t nvc = n_unit * (uint32_t)(NCH >> 1) + (uint32_t)(cidx >> 1);
              if (nvc > 0) PROF_WAIT(1, &bars[BAR_VFREE + 2 * wg + buf], (nvc - 1) & 1);
              TC_FENCE_AFTER();
            }
            tmem_st32(vb, hi);
            tmem_st32(vb + kPairK, lo);
            BACS_V(16, 3, 1) BACS_V(17, 3, 2) BACS_V(18, 3, 3) BACS_V(19, 3, 4)
            BACS_V(20, 4, 0) BACS_V(21, 4, 1) BACS_V(22, 4, 2) BACS_V(23, 4, 3)
            tmem_st16(vb + 32, hi);
            tmem_st16(vb + kPairK + 32, lo);
            BACS_V(24, 4, 4)
#undef BACS_V
#pragma unroll
            for (int k = 2; k < 8; ++k) hi[16 + k] = 0u, lo[16 + k] = 0u;
            tmem_st8(vb + 48, hi + 16);
            tmem_st8(vb + kPairK + 48, lo + 16);
            TC_WAIT_ST();
            TC_FENCE_BEFORE();
            mbar_arrive(&bars[BAR_VFULL + 2 * wg + buf]);
          }
        }
        // ------------------------------------------------ epilogue: row norms, rs -> tensor memory (A of the backward)
        PROF_MARK(8);   // forward phase (incl. its waits)
        PROF_WAIT(2, &bars[BAR_SFULL + wg], n_unit & 1);
        TC_FENCE_AFTER();
        {
          uint32_t d[32], r[32];
          tmem_ld32(tm + kColDS, d);
          TC_WAIT_LD();
#pragma unroll
          for (int y = 0; y < kRows; ++y) {
            const float S = __uint_as_float(d[y]) + __uint_as_float(d[kRows + y]);
            const float rs = S > 1e-37f ? rsqrt_fast_tc(S) : 0.f;  // zero row: norm 0, sub-gradient 0 (torch)
            if (active) loss_f = fmaf(S, rs, loss_f);
            const float rh = tf32_hi(rs);
            r[y] = __float_as_uint(rh);
            r[kRows + y] = __float_as_uint(rs - rh);
          }
          if (want_grad) {
            tmem_st32(tm + kColDS, r);
            TC_WAIT_ST();
            TC_FENCE_BEFORE();
            mbar_arrive(&bars[BAR_RFULL + wg]);
          }
        }
        ++n_unit;
        PROF_MARK(9);   // S wait + epilogue
        if (!want_grad) continue;
        // ------------------------------------------------ backward: chain rule on dV groups
        const bool last_blk = (blk == nblk - 1);
        float pend_own = 0.f, pend_low = 0.f;
        for (int j0 = 0; j0 < w; j0 += 8) {
          Corners cn;
          load_corners(j0, cn, true);
          float own8[8], low8[8];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int gq = ((j0 >> 1) + m) >> 1, buf = gq & 1;
            if ((m & 1) == 0) {
              PROF_WAIT(3, &bars[BAR_WREADY + 2 * wg + buf], ((n_unit - 1) * (uint32_t)(NG >> 1) + (uint32_t)(gq >> 1)) & 1);
              TC_FENCE_AFTER();
            }
            uint32_t wr[56];
            const uint32_t wb = tm + buf * kColVB + (m & 1) * kPairK;
            tmem_ld32(wb, wr);
            tmem_ld16(wb + 32, wr + 32);
            tmem_ld8(wb + 48, wr + 48);
            TC_WAIT_LD();
            if (m & 1) {
              TC_FENCE_BEFORE();
              mbar_arrive(&bars[BAR_WDONE + 2 * wg + buf]);
            }
            F2 Wv[25];
#pragma unroll
            for (int e = 0; e < 25; ++e) Wv[e] = pack32(wr[2 * e], wr[2 * e + 1]);
            F2 E[9], dE[9];
            make_E(cn, m, E);
            dE[0] = decoef<0, 0>(Wv, E); dE[1] = decoef<0, 1>(Wv, E); dE[2] = decoef<0, 2>(Wv, E);
            dE[3] = decoef<1, 0>(Wv, E); dE[4] = decoef<1, 1>(Wv, E); dE[5] = decoef<1, 2>(Wv, E);
            dE[6] = decoef<2, 0>(Wv, E); dE[7] = decoef<2, 1>(Wv, E); dE[8] = decoef<2, 2>(Wv, E);
            const int a = 2 * m;
            const F2 n00 = cn.n00[m], n01 = cn.n01[m], n10 = cn.n10[m], n11 = cn.n11[m];
            // d new_cd = sum_{c',d'} dE[(c+c')*3 + d+d'] * n_c'd'   (scaled by -2 coef at the row store)
            const F2 g00 = fma2(dE[4], n11, fma2(dE[3], n10, fma2(dE[1], n01, mul2(dE[0], n00))));
            const F2 g01 = fma2(dE[5], n11, fma2(dE[4], n10, fma2(dE[2], n01, mul2(dE[1], n00))));
            const F2 g10 = fma2(dE[7], n11, fma2(dE[6], n10, fma2(dE[4], n01, mul2(dE[3], n00))));
            const F2 g11 = fma2(dE[8], n11, fma2(dE[7], n10, fma2(dE[5], n01, mul2(dE[4], n00))));
            // columns: cell a owns (a, a+1), cell b = a+1 owns (a+1, a+2)
            own8[a] = pend_own + f2lo(g00);
            own8[a + 1] = f2lo(g01) + f2hi(g00);
            pend_own = f2hi(g01);
            low8[a] = pend_low + f2lo(g10);
            low8[a + 1] = f2lo(g11) + f2hi(g10);
            pend_low = f2hi(g11);
          }
          if (j0 + 8 >= w) {  // the last column is its own right neighbour
            own8[7] += pend_own;
            low8[7] += pend_low;
          }
          // intervals with several row blocks: partial rows wait in this thread's global scratch row
          if (nblk > 1) {
            float* so = scr + j0;
            float* sl = scr + (size_t)kChanCta * kMaxW + j0;
            if (blk > 0) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                own8[k] = so[k] + own8[k];
                low8[k] = sl[k] + low8[k];
              }
            }
            if (!last_blk) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                so[k] = own8[k];
                sl[k] = low8[k];
              }
            }
          }
          if (last_blk) {
            if (boundary_start) {
              float* bo = P.bnd_own + ((size_t)blockIdx.x * kChanCta + lrow) * kMaxW + j0;
#pragma unroll
              for (int k = 0; k < 8; ++k) bo[k] = own8[k];
            } else if (active) {
              float fin[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) fin[k] = gscale * (carry[j0 + k] + own8[k]);
              RowIO<T>::store8(reinterpret_cast<T*>(P.dnew) + (((size_t)b * P.A + ch) * h + i) * w + j0, fin);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) carry[j0 + k] = low8[k];
          }
        }
      }
      if (it == it1 - 1 && i < h - 1 && want_grad) {  // the interval below belongs to another CTA
        float* bl = P.bnd_low + ((size_t)blockIdx.x * kChanCta + lrow) * kMaxW;
        for (int j = 0; j < w; ++j) bl[j] = carry[j];
      }
      PROF_MARK(10);  // backward phase (incl. its waits)
      mbar_arrive(&bars[BAR_ITEM]);
    }
    if (warp == 0) PROF_DUMP(0);
    loss_d = (double)loss_f;
  } else if (warp < 11) {
    // =================================================== Bm operand builders (mask moments x row basis)
    const int aw = warp - 8;
    const int Yl = lane & 15, par = lane >> 4;
    uint32_t job = 0, unit = 0;
    PROF_DECL;
    const bool staged = P.fast_mask != 0;     // mask rows of a unit arrive by bulk copy one unit ahead; basis table in smem
    const int mpitch = W + 16;                // staged row pitch (bank spread)
    uint8_t* s_mrows = smem + sp.mrow_off;
    const float* s_phi = reinterpret_cast<const float*>(smem + sp.phi_off);
    // rows of the unit that starts at (item, block): image, first row, number of rows
    auto unit_rows = [&](int it, int blk, int& b, int& Y0, int& nr) {
      const int seg = it / h, i = it - seg * h;
      b = seg / P.ncb;
      Y0 = s_rowstart[i] + blk * kRows;
      nr = min(kRows, s_rowstart[i + 1] - Y0);
    };
    auto stage_unit = [&](int it, int blk, uint32_t u) {  // one elected lane: bulk copies of the unit's mask rows
      if (!staged || P.mask == nullptr || aw != 0 || lane != 0) return;
      int b, Y0, nr;
      unit_rows(it, blk, b, Y0, nr);
      uint64_t* bar = &bars[BAR_MASK + (u & 1)];
      mbar_expect_tx(bar, (uint32_t)(nr * W));
      uint8_t* dst = s_mrows + (size_t)(u & 1) * kRows * mpitch;
      for (int r = 0; r < nr; ++r)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + r * mpitch)),
                     "l"(P.mask + ((size_t)b * H + Y0 + r) * W), "r"((uint32_t)W), "r"(smem_u32(bar))
                     : "memory");
    };
    if (it0 < it1) stage_unit(it0, 0, 0);
    for (int it = it0; it < it1; ++it) {
      const int seg = it / h, i = it - seg * h, b = seg / P.ncb;
      const int Ybeg = s_rowstart[i], Yend = s_rowstart[i + 1];
      const int nblk = (Yend - Ybeg + kRows - 1) / kRows;
      for (int blk = 0; blk < nblk; ++blk, ++unit) {
        const int Y0 = Ybeg + blk * kRows, Y = Y0 + Yl;
        const bool rowok = Y < Yend;
        float py[5];
        bern4(rowok ? s_yw[Y] : 0.f, py);
        float* mom = s_mom + (((mom_buffers(w) == 2 ? (unit & 1) : 0) * kRows + Yl) * w) * 5;
        // the buffer of unit+1 was last read by unit-1, whose forward jobs every warp finished before the barrier below
        if (blk + 1 < nblk) stage_unit(it, blk + 1, unit + 1);
        else if (it + 1 < it1) stage_unit(it + 1, 0, unit + 1);
        if (staged && P.mask) PROF_WAIT(2, &bars[BAR_MASK + (unit & 1)], (unit >> 1) & 1);
        const uint8_t* mrow_s = s_mrows + ((size_t)(unit & 1) * kRows + Yl) * mpitch;
        for (int jl = 0; jl < jobs_per_unit; ++jl, ++job) {
          PROF_MARK(8);
          if (jl == NCH) asm volatile("bar.sync 1, 96;" ::: "memory");  // every moment of the unit is in shared memory
          PROF_MARK(4);
          if ((int)(job % 3) != aw) continue;
          const int slot = aw;              // job % kRing
          const uint32_t use = job / kRing;
          uint8_t* sl = smem + sp.ring_off + slot * kSlotBytes;
          if (jl < NCH) {
            // ---- forward chunk: cells (2c, 2c+1); this lane: row Yl, cell 2c + par ----
            const int cell = 2 * jl + par;
            float M[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (rowok && !(P.knock & 1)) {
              const int x0 = s_colstart[cell], x1 = s_colstart[cell + 1];
              if (staged) {
                // branch-free: M_k += m * phi_k[X], mask bytes from the staged row, basis values from the table
                if (P.mask == nullptr) {
                  for (int X = x0; X < x1; ++X) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) M[k] += s_phi[k * W + X];
                  }
                } else if (((x0 | (x1 - x0)) & 7) == 0) {
                  for (int X = x0; X < x1; X += 8) {
                    const uint2 v = *reinterpret_cast<const uint2*>(mrow_s + X);
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                      const uint32_t byte = ((t < 4 ? v.x : v.y) >> (8 * (t & 3))) & 0xffu;
                      const float mf = byte ? 1.f : 0.f;
#pragma unroll
                      for (int k = 0; k < 5; ++k) M[k] = fmaf(mf, s_phi[k * W + X + t], M[k]);
                    }
                  }
                } else {
                  for (int X = x0; X < x1; ++X) {
                    const float mf = mrow_s[X] ? 1.f : 0.f;
#pragma unroll
                    for (int k = 0; k < 5; ++k) M[k] = fmaf(mf, s_phi[k * W + X], M[k]);
                  }
                }
              } else {
                const uint8_t* mrow = P.mask ? P.mask + ((size_t)b * H + Y) * W : nullptr;
                auto add = [&](int X) {
                  float ph[5];
                  bern4(s_xw[X], ph);
#pragma unroll
                  for (int k = 0; k < 5; ++k) M[k] += ph[k];
                };
                if (mrow && ((reinterpret_cast<uintptr_t>(mrow + x0) & 7) == 0) && ((x1 - x0) & 7) == 0) {
                  for (int X = x0; X < x1; X += 8) {
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
              }
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) mom[cell * 5 + k] = M[k];
            // products of this lane's cell; slots 25..27 of a cell stay zero
            float v[28];
#pragma unroll
            for (int e = 0; e < 28; ++e) v[e] = e < 25 ? py[e / 5] * M[e % 5] : 0.f;
            if (use > 0) PROF_WAIT(1, &bars[BAR_BFREE + slot], (use - 1) & 1);   // the slot is written from here on
            PROF_MARK(8);
            // 16-byte groups [a_e, b_e, a_e+1, b_e+1] (e even): even groups are stored by the lane of cell a, odd
            // groups by the lane of cell b; the partner's pair of values comes by one exchange per value
#pragma unroll
            for (int g2 = 0; g2 < 7; ++g2) {
              const int Ga = 2 * g2, Gb = 2 * g2 + 1;  // group stored by par 0 / par 1
              // par 0 sends its values of group Gb and receives b's values of group Ga (and vice versa)
              const float s0 = par ? v[2 * Ga] : v[2 * Gb], s1 = par ? v[2 * Ga + 1] : v[2 * Gb + 1];
              const float r0 = __shfl_xor_sync(0xffffffffu, s0, 16), r1 = __shfl_xor_sync(0xffffffffu, s1, 16);
              const int G = par ? Gb : Ga;
              const float a0 = par ? r0 : v[2 * Ga], a1 = par ? r1 : v[2 * Ga + 1];
              const float b0 = par ? v[2 * Gb] : r0, b1 = par ? v[2 * Gb + 1] : r1;
              float4 hi4, lo4;
              hi4.x = tf32_hi(a0); hi4.y = tf32_hi(b0); hi4.z = tf32_hi(a1); hi4.w = tf32_hi(b1);
              lo4.x = a0 - hi4.x; lo4.y = b0 - hi4.y; lo4.z = a1 - hi4.z; lo4.w = b1 - hi4.w;
              uint8_t* dst = sl + G * 512 + (Yl >> 3) * 128 + (Yl & 7) * 16;
              *reinterpret_cast<float4*>(dst) = hi4;
              *reinterpret_cast<float4*>(dst + 256) = lo4;
            }
          } else {
            // ---- backward group: cells 4g .. 4g+3 as two pairs; rows along K ----
            const int g = jl - NCH;
            if (use > 0) PROF_WAIT(1, &bars[BAR_BFREE + slot], (use - 1) & 1);
            PROF_MARK(8);
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const int cell = 4 * g + 2 * pp + par;
              float M[5];
#pragma unroll
              for (int k = 0; k < 5; ++k) M[k] = mom[cell * 5 + k];
              uint8_t* base = sl + (Yl >> 2) * kBwdLbo + (Yl & 3) * 4;
#pragma unroll
              for (int e = 0; e < 28; ++e) {
                const float val = e < 25 ? py[e / 5] * M[e % 5] : 0.f;
                const float vh = tf32_hi(val);
                const int kap = pp * kPairK + 2 * e + par;
                uint8_t* dst = base + (kap >> 3) * 128 + (kap & 7) * 16;
                *reinterpret_cast<float*>(dst) = vh;
                *reinterpret_cast<float*>(dst + kBwdHalf) = val - vh;
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_BFULL + slot]);
          PROF_MARK(jl < NCH ? 9 : 10);   // build time of a forward / backward job
        }
        if (!want_grad) asm volatile("bar.sync 1, 96;" ::: "memory");   // forward-only: the unit's mask rows are consumed
      }
    }
    if (aw == 0) PROF_DUMP(1);
  } else {
    // =================================================== warp 11: MMA issue + TMA loads of the attention rows
    const bool leader = elect_one();
    const uint32_t ring = smem_u32(smem + sp.ring_off);
    PROF_DECL;
    uint32_t job = 0;
    uint32_t n_unit = 0, n_item_done = 0;
    constexpr uint32_t id32 = make_idesc(32), id16 = make_idesc(16), id112 = make_idesc(112);
    const uint32_t att_bytes = (uint32_t)(2 * P.box_chan * row_bytes);
    AttPlan ap{0, 0, -1, -1};
    auto wait_items = [&](uint32_t need) {  // phases of BAR_ITEM are consumed strictly one by one
      while (n_item_done < need) {
        PROF_WAIT(0, &bars[BAR_ITEM], n_item_done & 1);
        ++n_item_done;
      }
    };
    auto issue_att = [&](int slot, int seg, int row) {
      if (leader) {
        const int b = seg / P.ncb, cb = seg - b * P.ncb;
        uint8_t* dst = smem + sp.att_off + slot * sp.att_slot_bytes;
        mbar_expect_tx(&bars[BAR_ATT + slot], att_bytes);
        tma_load_3d(dst, &map_old, 0, row, b * P.A + cb * kChanCta, &bars[BAR_ATT + slot]);
        tma_load_3d(dst + kChanCta * row_bytes, &map_new, 0, row, b * P.A + cb * kChanCta, &bars[BAR_ATT + slot]);
      }
    };
    // rows of the first item
    if (it0 < it1) {
      const int seg = it0 / h, i = it0 - seg * h;
      ap = att_next(ap, true, i, h, P.n_att);
      issue_att(ap.slotU, seg, i);
      if (ap.newW >= 0) issue_att(ap.slotW, seg, i + 1);
    }
    for (int it = it0; it < it1; ++it) {
      const int seg = it / h, i = it - seg * h;
      // rows of the next item: into slots the current item does not read as soon as the previous item is done,
      // the others when the current item is done
      AttPlan nx = ap;
      int nseg = 0, ni = 0;
      bool have_next = it + 1 < it1;
      int deferU = -1, deferW = -1;
      if (have_next) {
        nseg = (it + 1) / h;
        ni = it + 1 - nseg * h;
        nx = att_next(ap, false, ni, h, P.n_att);
        wait_items((uint32_t)(it - it0));  // slots the current item does not read were last read by item it-1
        if (nx.newU >= 0) {
          if (nx.newU != ap.slotU && nx.newU != ap.slotW) issue_att(nx.newU, nseg, ni); else deferU = nx.newU;
        }
        if (nx.newW >= 0) {
          if (nx.newW != ap.slotU && nx.newW != ap.slotW) issue_att(nx.newW, nseg, ni + 1); else deferW = nx.newW;
        }
      }
      const int nrows = s_rowstart[i + 1] - s_rowstart[i];
      const int nblk = (nrows + kRows - 1) / kRows;
      for (int blk = 0; blk < nblk; ++blk) {
        // ---------------- forward: S[128 x 16] += V[128 x 8] * Bm^T per k-step; [hi|lo] rows of Bm as one N = 32 operand
        for (int c = 0; c < NCH; ++c, ++job) {
          const int slot = job % kRing;
          PROF_MARK(8);
          PROF_WAIT(1, &bars[BAR_BFULL + slot], (job / kRing) & 1);
          const uint32_t sb = ring + slot * kSlotBytes;
          // descriptor of k-step s: start address + s * 1024 bytes (the address field counts 16-byte units; the ring
          // lies below 256 KB, so the 14-bit field never carries)
          const uint64_t bd0 = make_desc(sb, 512, 128);
          const uint32_t bd_lo = (uint32_t)bd0, bd_hi = (uint32_t)(bd0 >> 32);
          for (int wg = 0; wg < 2; ++wg) {
            const int buf = c & 1;
            PROF_WAIT(2, &bars[BAR_VFULL + 2 * wg + buf], (n_unit * (uint32_t)(NCH >> 1) + (uint32_t)(c >> 1)) & 1);
            TC_FENCE_AFTER();
            PROF_MARK(11);
            if (leader) {
              const uint32_t tw = tmem + wg * kColsWg, va = tw + buf * kColVB;
              if (!(P.knock & 2))
#pragma unroll
              for (int s = 0; s < 7; ++s) {
                umma_tf32_ts2(tw + kColDS, va + 8 * s, bd_lo + s * 64, bd_hi, id32, (c | s) != 0);   // Vhi * [Bhi | Blo]
                // Vlo * Bhi joins the other small product: the main accumulator (columns 0..15) sees one addition per
                // k-step only, which halves the bias of the tensor core's truncating fp32 accumulation
                umma_tf32_ts2(tw + kColDS + kRows, va + kPairK + 8 * s, bd_lo + s * 64, bd_hi, id16, 1);
              }
              umma_commit(&bars[BAR_VFREE + 2 * wg + buf]);
              if (c == NCH - 1) umma_commit(&bars[BAR_SFULL + wg]);
            }
            PROF_MARK(12);   // issue of 14 MMAs + commits (leader lane)
            __syncwarp();
            PROF_MARK(13);
          }
          if (leader) umma_commit(&bars[BAR_BFREE + slot]);
          __syncwarp();
          PROF_MARK(14);
        }
        if (!want_grad) continue;
        // ---------------- backward: dV[128 x 112] = rs[128 x 16] * Bm[16 x 112], three split products
        for (int g = 0; g < NG; ++g, ++job) {
          const int slot = job % kRing;
          PROF_MARK(9);
          PROF_WAIT(3, &bars[BAR_BFULL + slot], (job / kRing) & 1);
          const uint32_t sb = ring + slot * kSlotBytes;
          const uint64_t gd0 = make_desc(sb, kBwdLbo, 128);
          const uint32_t gd_lo = (uint32_t)gd0, gd_hi = (uint32_t)(gd0 >> 32);
          for (int wg = 0; wg < 2; ++wg) {
            const int buf = g & 1;
            if (g == 0) PROF_WAIT(4, &bars[BAR_RFULL + wg], n_unit & 1);
            const uint32_t nwg = n_unit * (uint32_t)(NG >> 1) + (uint32_t)(g >> 1);   // group number on this buffer
            if (nwg > 0) PROF_WAIT(5, &bars[BAR_WDONE + 2 * wg + buf], (nwg - 1) & 1);
            TC_FENCE_AFTER();
            if (leader) {
              const uint32_t tw = tmem + wg * kColsWg, dv = tw + buf * kColVB, rs = tw + kColDS;
              if (!(P.knock & 2))
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const uint32_t bh = gd_lo + t * (2 * kBwdLbo / 16), bl = bh + kBwdHalf / 16;
                umma_tf32_ts2(dv, rs + 8 * t, bh, gd_hi, id112, t != 0);       // rs_hi * Bhi
                umma_tf32_ts2(dv, rs + 8 * t, bl, gd_hi, id112, 1);            // rs_hi * Blo
                umma_tf32_ts2(dv, rs + kRows + 8 * t, bh, gd_hi, id112, 1);    // rs_lo * Bhi
              }
              umma_commit(&bars[BAR_WREADY + 2 * wg + buf]);
            }
            __syncwarp();
          }
          if (leader) umma_commit(&bars[BAR_BFREE + slot]);
          __syncwarp();
        }
        ++n_unit;
      }
      if (have_next) {
        if (deferU >= 0 || deferW >= 0) {  // slots the current item still reads
          wait_items((uint32_t)(it - it0) + 1u);
          if (deferU >= 0) issue_att(deferU, nseg, ni);
          if (deferW >= 0) issue_att(deferW, nseg, ni + 1);
        }
        ap = nx;
      }
    }
    PROF_DUMP(2);
  }

  // ---- teardown: loss partial of the CTA, tensor memory ----
  TC_FENCE_BEFORE();
  const double tot = block_sum(loss_d, s_red);
  if (tid == 0) P.partials[blockIdx.x] = tot;
  __syncthreads();
  if (warp == 11) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// Completes the first gradient row of every CTA range that starts inside an image (low + own: two addends, order
// independent) and reduces the loss partials in a fixed order.  Grid (ranges, kFinishSlices): a block owns one slice of
// a range's 256 x w boundary row and every thread has its loads in flight together (the first version walked the row
// with one block per range, 32 dependent round trips to L2 per thread: 24 us cold / 10 us in the step).
constexpr int kFinishSlices = 8;
template <typename T>
__global__ void __launch_bounds__(256) distill_tc_finish_kernel(Params P, int ncta, double* __restrict__ loss_sum,
                                                               float* __restrict__ loss_scaled,
                                                               const float* __restrict__ addend) {
  __shared__ double scratch[32];
  const int k = blockIdx.x;
  pdl_wait();
  pdl_trigger();
  if (k == 0 && blockIdx.y == 0) {
    double s = 0.0;
    for (int i = threadIdx.x; i < ncta; i += blockDim.x) s += P.partials[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
      loss_sum[0] = s;
      if (loss_scaled) loss_scaled[0] = (float)((double)P.grad_coef * s) + (addend ? addend[0] : 0.f);
    }
  }
  if (!P.want_grad) return;
  const int it0 = P.bnd_info[2 * k], it1 = P.bnd_info[2 * k + 1];
  if (it0 >= it1) return;
  const int seg = it0 / P.h, i = it0 - seg * P.h;
  if (i == 0) return;
  int kp = k - 1;
  while (kp >= 0 && P.bnd_info[2 * kp] >= P.bnd_info[2 * kp + 1]) --kp;  // previous non-empty range
  if (kp < 0) return;
  const int b = seg / P.ncb, cb = seg - b * P.ncb;
  const float gscale = -2.f * P.grad_coef;
  const float* lowp = P.bnd_low + (size_t)kp * kChanCta * kMaxW;
  const float* ownp = P.bnd_own + (size_t)k * kChanCta * kMaxW;
  T* dn = reinterpret_cast<T*>(P.dnew);
  // slice blockIdx.y: channel rows [r0, r0 + 32) of the range's block; a thread owns columns j, j + 8, ... of one row
  constexpr int kRowsPerSlice = kChanCta / kFinishSlices, kPer = kRowsPerSlice * kMaxW / 256;
  const int r = (int)blockIdx.y * kRowsPerSlice + (threadIdx.x >> 3);
  const int ch = cb * kChanCta + r;
  if (ch >= P.A) return;
  float lo[kPer], ow[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int j = (threadIdx.x & 7) + 8 * q;
    lo[q] = j < P.w ? lowp[r * kMaxW + j] : 0.f;
    ow[q] = j < P.w ? ownp[r * kMaxW + j] : 0.f;
  }
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int j = (threadIdx.x & 7) + 8 * q;
    if (j < P.w) dn[(((size_t)b * P.A + ch) * P.h + i) * P.w + j] = DT<T>::from_f(gscale * (lo[q] + ow[q]));
  }
}

}  // namespace dtc
}  // namespace bacs

// ---------------------------------------------------------------- host side
namespace bacs {

typedef CUresult (*DtcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DtcEncodeTiledFn dtc_encode_fn() {
  static DtcEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DtcEncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

struct DtcConfig {
  int n_att, grid, fast_mask;
  size_t smem;
  size_t off_partials, off_own, off_low, off_info, off_scratch, ws_total;
};

// shape-only eligibility (the pointer alignment is checked at launch)
static bool dtc_config(int dtype, int B, int A, int h, int w, int H, int W, DtcConfig* cfg) {
  const int rb = w * (int)dtype_size(dtype);
  if (!(rb == 32 || rb == 64 || rb == 128) || w > dtc::kMaxW || (w % 8) != 0) return false;
  if (B <= 0 || A <= 0 || h <= 0 || H < h || W < w || H > 2048 || W > 2048) return false;
  if ((int64_t)B * A > 0x3fffffff || (int64_t)B * A * h > 0x3fffffff) return false;
  const size_t cap = 227 * 1024 - 1024 - 512;  // dynamic limit minus alignment slack and static shared memory
  bool ok = false;
  // preference: three attention-row slots (prefetch a whole item ahead), staged mask rows + basis table
  static const int tries[4][2] = {{3, 1}, {3, 0}, {2, 1}, {2, 0}};
  for (int t = 0; t < 4 && !ok; ++t) {
    if (tries[t][1] && (W % 16) != 0) continue;
    const dtc::SmemPlan sp = dtc::make_plan(tries[t][0], tries[t][1], rb, H, W, h, w);
    if ((size_t)sp.total <= cap) {
      cfg->n_att = tries[t][0];
      cfg->fast_mask = tries[t][1];
      cfg->smem = (size_t)sp.total + 1024;
      ok = true;
    }
  }
  if (!ok) return false;
  const int ncb = (A + dtc::kChanCta - 1) / dtc::kChanCta;
  const int64_t items = (int64_t)B * ncb * h;
  cfg->grid = (int)std::min<int64_t>(items, std::max(sm_count(), 1));
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  const size_t rows = (size_t)cfg->grid * dtc::kChanCta * dtc::kMaxW * sizeof(float);
  cfg->off_partials = take(sizeof(double) * cfg->grid);
  cfg->off_own = take(rows);
  cfg->off_low = take(rows);
  cfg->off_info = take(sizeof(int) * 2 * cfg->grid);
  cfg->off_scratch = take(2 * rows);
  cfg->ws_total = o;
  return true;
}

bool distill_tc_shape_ok(int dtype, int B, int A, int h, int w, int H, int W) {
  DtcConfig c;
  return dtc_encode_fn() != nullptr && dtc_config(dtype, B, A, h, w, H, W, &c);
}

size_t distill_tc_workspace_bytes(int dtype, int B, int A, int h, int w, int H, int W) {
  DtcConfig c;
  return dtc_config(dtype, B, A, h, w, H, W, &c) ? c.ws_total : 0;
}

// returns BACS_OK, an error, or +1 when the tensor-core path does not apply to these arguments (caller falls back)
int distill_tc_launch(const void* old_att, const void* new_att, int dtype, int B, int A, int h, int w, const uint8_t* mask, int H,
                      int W, float grad_coef, double* loss_sum, float* loss_scaled, const float* addend, void* dnew,
                      void* workspace, size_t workspace_bytes, cudaStream_t s) {
  DtcConfig cfg;
  DtcEncodeTiledFn enc = dtc_encode_fn();
  if (!enc || !dtc_config(dtype, B, A, h, w, H, W, &cfg)) return 1;
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!a16(old_att) || !a16(new_att) || (dnew && !a16(dnew)) || !a16(workspace)) return 1;
  if (workspace_bytes < cfg.ws_total) {
    set_error("bacs_teacher_distill: workspace %zu < %zu", workspace_bytes, cfg.ws_total);
    return BACS_ERR_WORKSPACE;
  }
  const size_t es = dtype_size(dtype);
  const int rb = w * (int)es;
  const int box_chan = (int)std::min<int64_t>(dtc::kChanCta, (int64_t)B * A);
  CUtensorMap maps[2];
  const CUtensorMapDataType dt = dtype == BACS_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == BACS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                      : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B * A};
  const cuuint64_t strides[2] = {(cuuint64_t)w * es, (cuuint64_t)h * w * es};
  const cuuint32_t box[3] = {(cuuint32_t)w, 1u, (cuuint32_t)box_chan};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const void* srcs[2] = {old_att, new_att};
  for (int k = 0; k < 2; ++k) {
    if (enc(&maps[k], dt, 3, const_cast<void*>(srcs[k]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 1;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  dtc::Params P;
  P.mask = mask;
  P.dnew = dnew;
  P.B = B; P.A = A; P.h = h; P.w = w; P.H = H; P.W = W;
  P.ncb = (A + dtc::kChanCta - 1) / dtc::kChanCta;
  P.total_items = B * P.ncb * h;
  P.sy = hp_scale(h, H);
  P.sx = hp_scale(w, W);
  P.grad_coef = grad_coef;
  P.partials = reinterpret_cast<double*>(ws + cfg.off_partials);
  P.bnd_own = reinterpret_cast<float*>(ws + cfg.off_own);
  P.bnd_low = reinterpret_cast<float*>(ws + cfg.off_low);
  P.bnd_info = reinterpret_cast<int*>(ws + cfg.off_info);
  P.mb_scratch = reinterpret_cast<float*>(ws + cfg.off_scratch);
  P.n_att = cfg.n_att;
  P.fast_mask = (cfg.fast_mask && (mask == nullptr || a16(mask))) ? 1 : 0;
  P.want_grad = dnew != nullptr;
  P.box_chan = box_chan;
  P.knock = getenv("BACS_DTC_KNOCK") ? atoi(getenv("BACS_DTC_KNOCK")) : 0;
#define BACS_DTC_LAUNCH(TT)                                                                                         \
  do {                                                                                                              \
    auto kern = dtc::distill_tc_kernel<TT>;                                                                         \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);        \
    if (e != cudaSuccess) {                                                                                         \
      set_error("bacs_teacher_distill: shared memory opt-in failed: %s", cudaGetErrorString(e));                   \
      return BACS_ERR_CUDA;                                                                                         \
    }                                                                                                               \
    launch_pdl(kern, dim3(cfg.grid), dim3(dtc::kThreads), cfg.smem, s, maps[0], maps[1], P);                        \
    BACS_CHECK_LAUNCH("bacs_teacher_distill(tensor cores)");                                                        \
    launch_pdl(dtc::distill_tc_finish_kernel<TT>, dim3(cfg.grid, dtc::kFinishSlices), dim3(256), 0, s, P,           \
               (int)cfg.grid, loss_sum, loss_scaled, addend);                                                       \
    BACS_CHECK_LAUNCH("bacs_teacher_distill(finish)");                                                              \
  } while (0)
  BACS_DISPATCH_DTYPE(dtype, TT, BACS_DTC_LAUNCH(TT));
#undef BACS_DTC_LAUNCH
#ifdef BACS_DTC_PROFILE
  {
    cudaStreamSynchronize(s);
    static long long prof[160 * 3 * 16];
    cudaMemcpyFromSymbol(prof, dtc::g_dtc_prof, sizeof(prof));
    const char* roles[3] = {"builder", "aux    ", "mma    "};
    for (int cta = 0; cta < std::min(cfg.grid, 148); cta += 49)
      for (int r = 0; r < 3; ++r) {
        fprintf(stderr, "cta %3d %s:", cta, roles[r]);
        for (int k = 0; k < 16; ++k) fprintf(stderr, " [%d]%lld", k, prof[(cta * 3 + r) * 16 + k]);
        fprintf(stderr, "\n");
      }
  }
#endif
#ifdef BACS_DTC_DEBUG
  {
    cudaStreamSynchronize(s);
    int dead = 0, trace[256];
    cudaMemcpyFromSymbol(&dead, dtc::g_dtc_dead, sizeof(int));
    cudaMemcpyFromSymbol(trace, dtc::g_dtc_trace, sizeof(trace));
    cudaMemcpyFromSymbol(&trace[255], dtc::g_dtc_barbase, sizeof(int));
    if (dead) {
      fprintf(stderr, "distill_tc DEBUG: %d waits timed out; barrier base offset %d\n", dead, trace[255]);
      for (int k = 0; k < std::min(dead, 63); ++k)
        if (trace[4 * k + 2] != 0)
        fprintf(stderr, "  cta %d thread %d (warp %d) barrier #%d parity %d\n", trace[4 * k], trace[4 * k + 1], trace[4 * k + 1] / 32,
                (trace[4 * k + 2] - trace[255]) / 8, trace[4 * k + 3]);
      const int zero = 0;
      cudaMemcpyToSymbol(dtc::g_dtc_dead, &zero, sizeof(int));
    }
  }
#endif
  return BACS_OK;
}

}  // namespace bacs
