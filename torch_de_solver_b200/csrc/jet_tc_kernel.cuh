// Fused Taylor-jet MLP loss + gradient kernel on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as jet_simt.cu for the interior (identity, K = 1) segments of nets whose hidden layers all have
// the same width W <= 104; the W x W layer GEMMs (forward, backward-data, weight gradient) run as
// tcgen05.mma.kind::tf32 with fp32 accumulators in TMEM, in the 3xTF32 split (hi*hi + hi*lo + lo*hi).
//
// Orientation: D[neuron, (point, channel)] = W . Y^T, i.e. accumulator lanes are neurons and columns are the
// (point, jet-channel) pairs of the tile (64 columns = 4 column parts of 16).  A thread that owns lane n sees
// every jet channel of its points for its neuron, so the tanh-jet rule and its adjoint are thread-local
// TMEM -> registers -> shared-memory epilogues, and what a thread writes for the next GEMM is a run of 16
// consecutive (point, channel) columns of its own row:
//   forward      D[n,(pc)]  = sum_k W[n,k]   Y[k,(pc)]      A = W image   (smem, K-major SW128)
//                                                           B = Y image   [k rows][(pc) contiguous], read MN-major
//   backward     D[k,(pc)]  = sum_n W^T[k,n] gZ[n,(pc)]     A = W^T image (smem, K-major SW128)
//                                                           B = gZ image  [n rows][(pc) contiguous], read MN-major
//   weight grad  dW[n,k]   += sum_pc gZ[n,(pc)] Y[k,(pc)]   A = gZ in TMEM (tcgen05.st from the epilogue registers)
//                                                           B = Y image   [k rows][(pc) contiguous], K-major SW128
// so every epilogue store is a 16-byte vector store.  MN-major tf32 operands use the 32-byte-base 128-byte swizzle
// (SWIZZLE_128B_BASE32B: 32-byte chunks XOR (row & 3)), K-major operands the canonical one (16-byte chunks XOR
// (row & 7)).  The dW accumulators stay in TMEM for the whole kernel and are flushed once per CTA; pre-activation
// jets are kept in registers between the forward and the backward sweep, nothing goes to HBM.
// Measured (profiles/microbench/mma_probe.cu): an M = 128 tf32 MMA with both operands in shared memory costs
// ~37 + N/4 cycles (the 4 KB A read dominates), with A in TMEM ~N/2 + 3; hence N = 64 and not less.
#pragma once
#include "common.cuh"

namespace tdb {

constexpr int kTcThreads = 512;                  // 16 warps: 4 lane windows x 4 column parts
constexpr int kTcParts = 4;
constexpr int kTcPC = 16;                        // columns per part
constexpr int kTcCols = kTcParts * kTcPC;        // (point, channel) columns per tile = MMA N
constexpr int kTcActBlock = 104 * 32;            // MN-major activation operand: [2 blocks of 32 columns][104 K rows][32]
constexpr int kTcActFloats = 2 * kTcActBlock;    // 6656 floats = 26 KB
constexpr int kTcYwRows = 112;                   // K-major operand of the weight-gradient GEMM: [2 blocks][112 rows][32]
constexpr int kTcYwBlock = kTcYwRows * 32;
constexpr int kTcYwFloats = 2 * kTcYwBlock;      // 7168 floats = 28 KB
constexpr int kTcMaxMma = 2;                     // W x W layers (their dW accumulators live in TMEM)
constexpr int kTcMaxOut = 4;                     // network outputs served by this kernel
// TMEM columns: D (forward / backward-data accumulator, two halves of 64 columns that the epilogue adds up) | gZ hi |
// gZ lo (A operands of the weight-gradient MMA) | dW slots (112 columns each)
constexpr uint32_t kTmD = 0, kTmAHi = 128, kTmALo = 192, kTmDw = 256, kTmDwCols = 112;


// float offset of element (K-row r, MN index k) of an MN-major tf32 operand (SWIZZLE_128B_BASE32B)
__host__ __device__ __forceinline__ int sw_off_mn(int r, int k, int rows) {
  return (k >> 5) * rows * 32 + r * 32 + ((((k & 31) >> 3) ^ (r & 3)) << 3) + (k & 7);
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                             uint32_t layout_type = 2 /* SWIZZLE_128B */) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;             // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// kind::tf32 instruction descriptor: D = f32, A = B = tf32, majors, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");   // suspend-time hint: sleep, do not spin
  } while (!done);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// issue only: pair with tmem_ld_wait() (several loads in flight behind one wait)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st1(uint32_t taddr, float a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(__float_as_uint(a)) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                  "r"(__float_as_uint(v[3])) : "memory");
}
// store `count` (1..8, warp-uniform) consecutive columns
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const float* v, int count) {
  int c = 0;
  if (count & 4) { tmem_st4(taddr, v); c = 4; }
  if (count == 8) { tmem_st4(taddr + 4, v + 4); return; }
  if (count & 2) { tmem_st2(taddr + c, v + c); c += 2; }
  if (count & 1) tmem_st1(taddr + c, v[c]);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split_store(float* hi_buf, float* lo_buf, int off, float y) {
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(y));
  const float h = __uint_as_float(hb);
  hi_buf[off] = h;
  lo_buf[off] = y - h;
}


__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
         "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
         "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
         "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
         "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}

// tf32 value nearest to x, ties away from zero: the result of cvt.rna.tf32.f32 for every finite x, computed on the
// integer ALU (the conversion instruction runs on the quarter-rate pipe and its latency showed up as the top stall of
// the epilogues: 32 conversions per thread and layer)
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split16(const float* v, float* hi, float* lo) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[j] = tf32_rna(v[j]);
    lo[j] = v[j] - hi[j];
  }
}
__device__ __forceinline__ void st4(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------------------------
// MMA issue helpers.  Called by every lane of one warp (warp-uniform control flow keeps the descriptors in uniform
// registers); each tcgen05.mma itself is issued by the elected lane.  All operand buffers are 1024-byte aligned.
// ------------------------------------------------------------------------------------------------
// 3xTF32 GEMM of one layer: A = weight image (K-major SW128), B = activation image [K rows][(pc) columns] (MN-major
// BASE32B), hi image followed by lo image = 4 column blocks of 32.  Two MMAs per K-step instead of three:
//   D[:, 0:128]  (+)= W_hi . [Y_hi | Y_lo]      (N = 128: both products that share the A operand in one instruction -
//                                                 an SS-mode MMA costs ~37 + N/4 cycles, the A read dominates)
//   D[:, 0:64]    +=  W_lo . Y_hi               (N = 64)
// and the epilogue adds the two halves: Z = D[:, 0:64] + D[:, 64:128].
template <int KS>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, const float* w_hi, const float* w_lo,
                                           const float* b_hi, const float* /*b_lo: contiguous after b_hi*/) {
  constexpr uint32_t idesc128 = umma_idesc(128, 2 * kTcCols, 0, 1), idesc64 = umma_idesc(128, kTcCols, 0, 1);
  const uint64_t dwh = umma_desc(smem_u32(w_hi), 16, 1024), dwl = umma_desc(smem_u32(w_lo), 16, 1024);
  // MN blocks of 32 columns at LBO = one block; 8 K rows = two 4-row swizzle atoms (SBO = 512 B)
  const uint64_t dbh = umma_desc(smem_u32(b_hi), kTcActBlock * 4, 512, 1);
  const bool leader = elect_one();
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    const uint64_t ao = ((uint64_t)(s >> 2) * kTcWBlock * 4 + (uint64_t)(s & 3) * 32) >> 4;
    const uint64_t bo = ((uint64_t)s * 1024) >> 4;
    if (leader) {
      umma_tf32(d_tmem, dwh + ao, dbh + bo, idesc128, s ? 1u : 0u);
      umma_tf32(d_tmem, dwl + ao, dbh + bo, idesc64, 1u);
    }
  }
}
__device__ __forceinline__ void issue_gemm_any(uint32_t d_tmem, const float* w_hi, const float* w_lo,
                                               const float* b_hi, const float* b_lo, int ksteps) {
  if (ksteps <= 4) issue_gemm<4>(d_tmem, w_hi, w_lo, b_hi, b_lo);             // zero-padded images: extra steps add 0
  else if (ksteps <= 8) issue_gemm<8>(d_tmem, w_hi, w_lo, b_hi, b_lo);
  else issue_gemm<13>(d_tmem, w_hi, w_lo, b_hi, b_lo);
}
// dW[128 (n) x 112 (k)] += A(gZ in TMEM: lanes n, columns (pc)) . B(Y image [k rows][64 columns], K-major SW128):
// MMAs [I0, I1) of the 24 of one weight-gradient GEMM (index = pass * 8 + K-step).  The issue of a tcgen05.mma
// blocks while the (short) MMA queue is full, so the issuing warp feeds the 24 MMAs in small portions between the
// pieces of its own epilogue instead of falling ~1400 cycles behind the other warps.
template <int I0, int I1>
__device__ __forceinline__ void issue_wgrad_range(uint32_t d_tmem, uint32_t a_hi_tmem, uint32_t a_lo_tmem,
                                                  const float* y_hi, const float* y_lo, uint32_t accumulate) {
  if (I0 >= I1) return;
  constexpr uint32_t idesc = umma_idesc(128, 112, 0, 0);
  const uint64_t dyh = umma_desc(smem_u32(y_hi), 16, 1024), dyl = umma_desc(smem_u32(y_lo), 16, 1024);
  const bool leader = elect_one();
#pragma unroll
  for (int i = I0; i < I1; ++i) {
    const int pass = i >> 3, s = i & 7;
    const uint32_t A = pass == 0 ? a_lo_tmem : a_hi_tmem;
    const uint64_t B = pass == 1 ? dyl : dyh;
    const uint64_t bo = ((uint64_t)(s >> 2) * kTcYwBlock * 4 + (uint64_t)(s & 3) * 32) >> 4;
    if (leader) umma_tf32_ts(d_tmem, A + (uint32_t)s * 8, B + bo, idesc, i ? 1u : accumulate);
  }
}

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
// tanh(x) = 1 - 2 / (exp(2x) + 1): two MUFU ops, absolute error ~1e-7 over the whole range (the jets only ever
// use tanh through a, 1 - a^2, ...: absolute, not relative, accuracy is what the residual sees)
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

// warp-wide sums of 16 per-lane values: afterwards the lane pair (2q, 2q + 1) holds the total of column
// c(q) = bit-reversed-ish index 8*b4 + 4*b3 + 2*b2 + b1 of the lane number (16 shuffles)
__device__ __forceinline__ float warp_multi_reduce16(float* v, int lane) {
#pragma unroll
  for (int off = 16, cnt = 8; off >= 2; off >>= 1, cnt >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const float keep = up ? v[i + cnt] : v[i];
      const float send = up ? v[i] : v[i + cnt];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int reduce16_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// bulk (TMA, 1-D) copy of one weight image pair into shared memory, completion on an mbarrier
__device__ __forceinline__ void bulk_load_image(float* dst, const float* src, uint64_t* bar) {
  constexpr uint32_t kBytes = 2 * kTcWFloats * 4, kChunk = 8192;
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(kBytes) : "memory");
  for (uint32_t o = 0; o < kBytes; o += kChunk)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst) + o), "l"(reinterpret_cast<const char*>(src) + o), "r"(kChunk), "r"(b)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
constexpr int kTcMaxTerms = 48, kTcMaxFactors = 96;   // operator program cached in shared memory
constexpr int kTcMaxPts = 32;                         // points per tile (J = 2)
// float offsets of the big buffers from the 1024-byte aligned base (compile-time: addresses fold into immediates)
constexpr int kOffWHi = 0, kOffWLo = kTcWFloats, kOffActHi = 2 * kTcWFloats, kOffActLo = kOffActHi + kTcActFloats,
              kOffYwHi = kOffActLo + kTcActFloats, kOffYwLo = kOffYwHi + kTcYwFloats, kOffX = kOffYwLo + kTcYwFloats,
              kOffU = kOffX + 2 * kTcMaxPts * 4, kOffGu = kOffU + kTcMaxOut * kTcCols, kOffUP = kOffGu + kTcMaxOut * kTcCols,
              kOffCg = kOffUP + 4 * kTcMaxOut * kTcCols, kOffBl = kOffCg + (kMaxCParams + 3) / 4 * 4,
              kOffEnd = kOffBl + kTcMaxOut;                      // kOffBl: last-layer bias
struct TcSmem {
  tdb200_term* termS;
  tdb200_factor* facS;
  tdb200_segment* segS;
  float* scaleS;          // [32] lambda / len per slot
  double* lossT;          // [kTcMaxPts][TDB200_MAX_COLS] per-point-thread loss accumulators (no atomics)
  int4* recS;             // [kTcMaxTerms] pre-decoded terms (<= 2 live factors, integer powers <= 3), see below
  int* fastS;             // 1: every term of the segment has a record
  uint64_t *bar, *wbar, *gbar;
  uint32_t* tmem_ptr;
};
constexpr size_t kTcSmemBytes =
    (size_t)(2 * kTcWFloats + 2 * kTcActFloats + 2 * kTcYwFloats) * 4 + 1024 /*align*/ +
    (2 * kTcMaxPts * 4 + 2 * kTcMaxOut * kTcCols + 4 * kTcMaxOut * kTcCols + kMaxCParams + kTcMaxOut) * 4 + 64 + 64 +
    kTcMaxTerms * sizeof(tdb200_term) + kTcMaxFactors * sizeof(tdb200_factor) + 16 + sizeof(tdb200_segment) + 32 * 4 +
    kTcMaxPts * TDB200_MAX_COLS * 8 + 64 + kTcMaxTerms * 16 + 16;


// Jet signature: derivative orders of up to three directions (sorted by input axis); J = 1 + O0 + O1 + O2.
template <int O0, int O1, int O2, int NMMA>
__global__ void __launch_bounds__(kTcThreads, 1) jet_tc_kernel(const JetArgs a, const float* __restrict__ wimg) {
  constexpr int J = 1 + O0 + O1 + O2;
  constexpr int ND = (O0 > 0) + (O1 > 0) + (O2 > 0);
  constexpr int PH = kTcPC / J > kTcMaxPts / kTcParts ? kTcMaxPts / kTcParts : kTcPC / J;   // points per column part
                                               // (J = 1: capped by the per-tile point tables, 8 of 16 columns used)
  constexpr int P = kTcParts * PH;             // points per tile
  constexpr int C = PH * J;                    // used columns per part (<= 16)
  constexpr int JD = J > 1 ? J - 1 : 1;        // derivative channels per point
  constexpr int ORD[3] = {O0, O1, O2};
  extern __shared__ uint8_t smem_raw_tc[];
  // align inside the shared window with pointer arithmetic on the __shared__ symbol itself: a round trip through
  // uintptr_t would turn every later access into a generic LD / ST
  const uint32_t s0_ = smem_u32(smem_raw_tc);
  float* const sbase = reinterpret_cast<float*>(smem_raw_tc + (((s0_ + 1023u) & ~1023u) - s0_));
  TcSmem sm;
  {
    // kOffEnd is a multiple of 4 floats and sbase is 1024-byte aligned: everything below is 16-byte aligned
    uint8_t* q = reinterpret_cast<uint8_t*>(sbase + kOffEnd);
    sm.bar = reinterpret_cast<uint64_t*>(q); q += 32;
    sm.wbar = sm.bar + 1;
    sm.gbar = sm.bar + 2;
    sm.tmem_ptr = reinterpret_cast<uint32_t*>(sm.bar + 3);
    sm.termS = reinterpret_cast<tdb200_term*>(q); q += (kTcMaxTerms * sizeof(tdb200_term) + 15) / 16 * 16;
    sm.facS = reinterpret_cast<tdb200_factor*>(q); q += (kTcMaxFactors * sizeof(tdb200_factor) + 15) / 16 * 16;
    sm.segS = reinterpret_cast<tdb200_segment*>(q); q += (sizeof(tdb200_segment) + 15) / 16 * 16;
    sm.scaleS = reinterpret_cast<float*>(q); q += 32 * 4;
    sm.lossT = reinterpret_cast<double*>(q); q += kTcMaxPts * TDB200_MAX_COLS * 8;
    sm.recS = reinterpret_cast<int4*>(q); q += kTcMaxTerms * 16;
    sm.fastS = reinterpret_cast<int*>(q);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = (warp & 3) * 32 + lane;                 // neuron = TMEM lane owned by this thread
  const int part = warp >> 2;                           // which column part (0..3) of the tile this thread owns
  const int L = a.n_layers, W = a.widths[1], n_out = a.widths[L], d = a.d;
  const int ksteps = (W + 7) / 8;
  const bool live = n < W;
  const int col0 = part * kTcPC;                        // first (point, channel) column of this thread
  // one gradient-partial row per CTA: the four column parts are added up in shared memory in a fixed order at the end
  // (a single owner thread per address -> bit-reproducible; the reduction kernel reads 4x fewer rows)
  float* const my_grad = a.part_grad + (size_t)blockIdx.x * a.n_params_pad;
  // where this thread's 16 columns live in the operand images (floats)
  const int actA = (part >> 1) * kTcActBlock + n * 32 + ((((part & 1) * 2) ^ (n & 3)) << 3);       // columns 0..7
  const int actB = (part >> 1) * kTcActBlock + n * 32 + ((((part & 1) * 2 + 1) ^ (n & 3)) << 3);   // columns 8..15
  int ywq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) ywq[q] = (part >> 1) * kTcYwBlock + n * 32 + ((((part & 1) * 4 + q) ^ (n & 7)) << 2);

  // ---- one-time setup --------------------------------------------------------------------------------
  for (int i = tid; i < a.n_params_pad; i += kTcThreads) a.part_grad[(size_t)blockIdx.x * a.n_params_pad + i] = 0.f;
  for (int i = tid; i < 2 * kTcActFloats + 2 * kTcYwFloats; i += kTcThreads) (sbase + kOffActHi)[i] = 0.f;   // pad rows / columns stay zero
  if (tid < kMaxCParams) (sbase + kOffCg)[tid] = 0.f;
  for (int i = tid; i < 2 * kTcMaxPts * 4; i += kTcThreads) (sbase + kOffX)[i] = 0.f;      // axes >= d stay zero
  for (int i = tid; i < kTcMaxOut * kTcCols; i += kTcThreads) (sbase + kOffGu)[i] = 0.f;   // pad columns stay zero
  if (tid < kTcMaxOut) (sbase + kOffBl)[tid] = tid < n_out ? a.arena[a.b_off[L - 1] + tid] : 0.f;
  for (int i = tid; i < min(kTcMaxTerms, a.n_terms); i += kTcThreads) sm.termS[i] = a.terms[i];
  for (int i = tid; i < min(kTcMaxFactors, a.n_factors); i += kTcThreads) sm.facS[i] = a.factors[i];
  for (int i = tid; i < (int)(sizeof(tdb200_segment) / 4); i += kTcThreads)
    reinterpret_cast<uint32_t*>(sm.segS)[i] = reinterpret_cast<const uint32_t*>(a.segs)[i];
  if (tid < a.n_slots) sm.scaleS[tid] = a.slot_scale[tid];
  for (int i = tid; i < kTcMaxPts * TDB200_MAX_COLS; i += kTcThreads) sm.lossT[i] = 0.0;
  if (tid == 0) { mbar_init(sm.bar, 1); mbar_init(sm.wbar, 1); mbar_init(sm.gbar, 1); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(sm.tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sm.tmem_ptr;
  // pre-decode the operator program: term -> {coeff bits | buffer index, kind, u offsets of <= 2 live factors, powers}
  if (tid == 0) *sm.fastS = 1;
  __syncthreads();
  if (tid < min(kTcMaxTerms, a.n_terms) && tid < sm.segS->col_term_end[sm.segS->n_cols - 1]) {
    const tdb200_term tm = sm.termS[tid];
    int off[2] = {0xFFFF, 0xFFFF}, ipw[2] = {0, 0}, nf = 0;
    bool ok = tm.kind == 0 || (tm.idx >= 0 && tm.idx < 0x7fffffffLL);
    for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
      const tdb200_factor fc = sm.facS[fi];
      if (fc.ipow == 0) continue;                       // x^0: contributes 1 and no derivative
      if (fc.ipow < 0 || fc.ipow > 3 || nf == 2) { ok = false; break; }
      off[nf] = fc.var * kTcCols + fc.chan;
      ipw[nf] = fc.ipow;
      ++nf;
    }
    if (ok) sm.recS[tid] = make_int4(tm.kind == 0 ? __float_as_int(tm.coeff) : (int)tm.idx, tm.kind, off[0] | (off[1] << 16),
                                     ipw[0] | (ipw[1] << 8));
    else *sm.fastS = 0;
  }
  __syncthreads();
  const bool fast_op = *sm.fastS != 0;
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);    // this warp's lane window
#ifdef TDB_TC_TIMING
  long long tacc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tacc[i] = 0;
  long long tlast = clock64();
#endif
#ifdef TDB_TC_TIMING       // phase timers: build with TDB200_TC_TIMING_BUILD=1 (adds ~6 % instructions)
#define TMARK(i) do { if (a.dbg) { const long long tn_ = clock64(); tacc[i] += tn_ - tlast; tlast = tn_; } } while (0)
#else
#define TMARK(i) do { } while (0)
#endif
  uint32_t phase = 0, wphase = 0, gphase = 0;           // wphase is only used by warp 0
  bool wgrad_pending = false;                           // weight-gradient MMAs still reading TMEM A / the Y image
  uint32_t dw_started = 0;
  // per-layer parameters this thread needs all the time, and its gradient accumulators (flushed once per CTA)
  float bias[NMMA + 1], w0[4] = {0.f, 0.f, 0.f, 0.f}, wl[kTcMaxOut];
  float db_acc[NMMA + 1], dw0_acc[4] = {0.f, 0.f, 0.f, 0.f}, dw0_dir[3] = {0.f, 0.f, 0.f}, dwl_acc[kTcMaxOut], dbl_acc = 0.f;
#pragma unroll
  for (int l = 0; l <= NMMA; ++l) { bias[l] = live ? a.arena[a.b_off[l] + n] : 0.f; db_acc[l] = 0.f; }
  if (live)
    for (int ax = 0; ax < d; ++ax) w0[ax] = a.arena[a.w_off[0] + n * d + ax];
#pragma unroll
  for (int v = 0; v < kTcMaxOut; ++v) { wl[v] = (live && v < n_out) ? a.arena[a.w_off[L - 1] + v * W + n] : 0.f; dwl_acc[v] = 0.f; }

  const tdb200_segment& sg = *sm.segS;                  // shared-memory copy (set up above, visible after the sync)
  const int ncols = sg.n_cols;
  float w0d[3];                                         // first-layer weight along each jet direction
#pragma unroll
  for (int i = 0; i < 3; ++i)
    w0d[i] = i < ND ? fmaf(w0[0], sg.dir_vec[i][0], fmaf(w0[1], sg.dir_vec[i][1], fmaf(w0[2], sg.dir_vec[i][2], w0[3] * sg.dir_vec[i][3])))
                      : 0.f;     // first-order input of direction i: W0[n, :] . v_i (column of W0 for a pure partial)

  if (tid == 0) bulk_load_image((sbase + kOffWHi), wimg, sm.wbar);          // W_1 for the first tile

  auto load_points = [&](int tile_idx, float* dst) {
    const long long gf = (long long)tile_idx * P;
    const int pv = (int)min((long long)P, sg.n_groups - gf);
    for (int i = tid; i < P * d; i += kTcThreads) {
      const int p = i / d, ax = i - p * d;
      if (p < pv)      // asynchronous copy: nobody waits for the load until the next tile starts
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst + p * 4 + ax)),
                     "l"(a.pts + (size_t)(sg.pts_off + gf + p) * d + ax) : "memory");
      else
        dst[p * 4 + ax] = 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // Y_l of this thread's columns from the saved tanh values and pre-activation jets
  auto jets_from_saved = [&](const float* as_l, const float (*zd_l)[JD], bool first, float* y) {
#pragma unroll
    for (int p = 0; p < PH; ++p) {
      const TanhF f(as_l[p]);
      y[p * J] = as_l[p];
      int c = 1;
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        float z[4] = {0.f, 0.f, 0.f, 0.f}, yy[4];
        if (first) z[0] = w0d[i];
        else {
#pragma unroll
          for (int k = 0; k < ORD[i]; ++k) z[k] = zd_l[p][c - 1 + k];
        }
        tanh_jet_fwd(f, z, ORD[i], yy);
#pragma unroll
        for (int k = 0; k < ORD[i]; ++k) y[p * J + c + k] = yy[k];
        c += ORD[i];
      }
    }
#pragma unroll
    for (int j = C; j < 16; ++j) y[j] = 0.f;
  };
  auto store_act = [&](const float* v) {                // 16 columns -> MN-major operand image (hi / lo)
    float hi[16], lo[16];
    split16(v, hi, lo);
    st4((sbase + kOffActHi) + actA, hi); st4((sbase + kOffActHi) + actA + 4, hi + 4); st4((sbase + kOffActHi) + actB, hi + 8); st4((sbase + kOffActHi) + actB + 4, hi + 12);
    st4((sbase + kOffActLo) + actA, lo); st4((sbase + kOffActLo) + actA + 4, lo + 4); st4((sbase + kOffActLo) + actB, lo + 8); st4((sbase + kOffActLo) + actB + 4, lo + 12);
  };

  load_points(blockIdx.x, (sbase + kOffX));
  int xbuf = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const long long g_first = (long long)tile * P;
    const int p_valid = (int)min((long long)P, sg.n_groups - g_first);
    float* const xcur = (sbase + kOffX) + xbuf * kTcMaxPts * 4;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                    // this tile's points (loaded one tile ahead) are visible
    if (tile + (int)gridDim.x < a.n_tiles) load_points(tile + gridDim.x, (sbase + kOffX) + (xbuf ^ 1) * kTcMaxPts * 4);
    xbuf ^= 1;
    if (tid < p_valid && sg.tgt_off >= 0)               // warm L1 for the operator phase of this tile
      asm volatile("prefetch.global.L1 [%0];" :: "l"(a.targets + sg.tgt_off + (g_first + tid) * ncols) : "memory");
    TMARK(0);

    float as[NMMA + 1][PH];                             // tanh values of layers 0..NMMA for this thread's points
    float zd[NMMA > 0 ? NMMA : 1][PH][JD];              // pre-activation derivative channels of layers 1..NMMA
    float y[16];

    // ---- layer 0 (K = d): thread-local ----------------------------------------------------------------
#pragma unroll
    for (int p = 0; p < PH; ++p) {
      const float4 x4 = *reinterpret_cast<const float4*>(xcur + (part * PH + p) * 4);   // unused axes hold 0
      as[0][p] = tanh_fast(fmaf(w0[0], x4.x, fmaf(w0[1], x4.y, fmaf(w0[2], x4.z, fmaf(w0[3], x4.w, bias[0])))));
    }
    jets_from_saved(as[0], nullptr, true, y);
    if (live) store_act(y);

    // ---- W x W layers: tensor-core GEMM + thread-local tanh-jet epilogue ------------------------------
#pragma unroll
    for (int l = 1; l <= NMMA; ++l) {
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      TMARK(1);
      if (warp == 0) {
        mbar_wait(sm.wbar, wphase); wphase ^= 1;        // W_l image has landed
        tc_fence_after();
        issue_gemm_any(tmem + kTmD, (sbase + kOffWHi), (sbase + kOffWLo), (sbase + kOffActHi), (sbase + kOffActLo), ksteps);
        if (elect_one()) umma_commit(sm.bar);
        __syncwarp();
        TMARK(2);
      }
      mbar_wait(sm.bar, phase);
      phase ^= 1;
      tc_fence_after();
      TMARK(3);
      // next image (W_{l+1}, or W_NMMA^T for the backward sweep, or W_1 again) streams in behind the epilogue
      if (tid == 0) {
        const float* nxt = l < NMMA ? wimg + (size_t)l * 4 * kTcWFloats
                                    : (a.do_grad ? wimg + (size_t)(NMMA - 1) * 4 * kTcWFloats + 2 * kTcWFloats : wimg);
        bulk_load_image((sbase + kOffWHi), nxt, sm.wbar);
      }
      float z[16];
      {
        float z2[16];
        tmem_ld16(t_lane + kTmD + (uint32_t)col0, z);
        tmem_ld16(t_lane + kTmD + (uint32_t)(kTcCols + col0), z2);
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] += z2[j];
      }
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        as[l][p] = tanh_fast(z[p * J] + bias[l]);
#pragma unroll
        for (int k = 0; k < J - 1; ++k) zd[l - 1][p][k] = z[p * J + 1 + k];
      }
      jets_from_saved(as[l], zd[l - 1], false, y);
      if (live && l < NMMA) store_act(y);
    }

    TMARK(4);
    // ---- last layer: u[v][col] = sum_n Wl[v][n] y[n][col]  (warp multi-value reduction, fixed order) ------
    for (int v = 0; v < n_out; ++v) {
      float t16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) t16[j] = wl[v] * y[j];          // wl is zero in dead lanes
      const float tot = warp_multi_reduce16(t16, lane);
      if ((lane & 1) == 0) (sbase + kOffUP)[((warp & 3) * kTcMaxOut + v) * kTcCols + col0 + reduce16_col(lane)] = tot;
    }
    __syncthreads();
    for (int idx = tid; idx < n_out * kTcCols; idx += kTcThreads) {
      const int v = idx / kTcCols, r = idx - v * kTcCols;
      const int jc = r & (kTcPC - 1);
      float s = (jc < C && jc % J == 0) ? (sbase + kOffBl)[v] : 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) s += (sbase + kOffUP)[(w * kTcMaxOut + v) * kTcCols + r];
      (sbase + kOffU)[idx] = s;
      (sbase + kOffGu)[idx] = 0.f;
    }
    __syncthreads();

    TMARK(5);
    // ---- operator terms, residual, loss, adjoint seeds (one thread per point) -------------------------
    if (tid < p_valid && fast_op) {
      // pre-decoded terms: no factor loops, powers by selection
      const int p = tid;
      const int pc = (p / PH) * kTcPC + (p % PH) * J;   // first column of this point
      const long long row = g_first + p;
      const float* u = (sbase + kOffU) + pc;
      float* gu = (sbase + kOffGu) + pc;
      auto pw = [](float x, int i) { const float x2 = x * x; return i == 1 ? x : i == 2 ? x2 : i == 3 ? x2 * x : 1.f; };
      auto dpw = [](float x, int i) { return i == 1 ? 1.f : i == 2 ? 2.f * x : i == 3 ? 3.f * x * x : 0.f; };
      for (int col = 0; col < ncols; ++col) {
        const int tb = sg.col_term_begin[col], te = sg.col_term_end[col];
        float val = 0.f;
        for (int t = tb; t < te; ++t) {
          const int4 r = sm.recS[t];
          const float cf = r.y == 0 ? __int_as_float(r.x) : r.y == 1 ? __ldg(a.coeffs + r.x + row) : a.arena[a.n_net_params + r.x];
          const int o0 = r.z & 0xFFFF, o1 = (r.z >> 16) & 0xFFFF;
          const float x0 = o0 != 0xFFFF ? u[o0] : 1.f, x1 = o1 != 0xFFFF ? u[o1] : 1.f;
          val = fmaf(cf * pw(x0, r.w & 255), pw(x1, r.w >> 8), val);
        }
        if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
        const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
        const float res = val - tgt;
        const float rw = a.row_weight ? __ldg(a.row_weight + row) : 1.f;        // causal-loss weight (no grad)
        sm.lossT[p * TDB200_MAX_COLS + col] += (double)rw * (double)res * (double)res;
        if (!a.do_grad) continue;
        const float seed = a.field_seed ? __ldg(a.field_seed + sg.field_off + row * ncols + col)
                                        : 2.f * sm.scaleS[sg.col_slot[col]] * rw * res;
        for (int t = tb; t < te; ++t) {
          const int4 r = sm.recS[t];
          const float cf = r.y == 0 ? __int_as_float(r.x) : r.y == 1 ? __ldg(a.coeffs + r.x + row) : a.arena[a.n_net_params + r.x];
          const int o0 = r.z & 0xFFFF, o1 = (r.z >> 16) & 0xFFFF;
          const float x0 = o0 != 0xFFFF ? u[o0] : 1.f, x1 = o1 != 0xFFFF ? u[o1] : 1.f;
          const float p0 = pw(x0, r.w & 255), p1 = pw(x1, r.w >> 8), sc = seed * cf;
          if (o0 != 0xFFFF) gu[o0] += sc * dpw(x0, r.w & 255) * p1;
          if (o1 != 0xFFFF) gu[o1] += sc * p0 * dpw(x1, r.w >> 8);
          if (r.y == 2) atomicAdd(&(sbase + kOffCg)[r.x], seed * p0 * p1);
        }
      }
    } else     if (tid < p_valid) {
      const int p = tid;
      const int pc = (p / PH) * kTcPC + (p % PH) * J;   // first column of this point
      const long long row = g_first + p;
      for (int col = 0; col < ncols; ++col) {
        float val = 0.f;
        for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
          const tdb200_term tm = sm.termS[t];
          float prod = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                               : a.arena[a.n_net_params + tm.idx];
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = sm.facS[fi];
            prod *= pow_i((sbase + kOffU)[fc.var * kTcCols + pc + fc.chan], fc.ipow, fc.pow);
          }
          val += prod;
        }
        TMARK(11);
        if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
        const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
        const float res = val - tgt;
        const int slot = sg.col_slot[col];
        sm.lossT[p * TDB200_MAX_COLS + col] += (double)res * (double)res;
        TMARK(12);
        if (!a.do_grad) continue;
        const float seed = a.field_seed ? __ldg(a.field_seed + sg.field_off + row * ncols + col) : 2.f * sm.scaleS[slot] * res;
        for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
          const tdb200_term tm = sm.termS[t];
          const float cf = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                                   : a.arena[a.n_net_params + tm.idx];
          float full = 1.f;
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = sm.facS[fi];
            const float x = (sbase + kOffU)[fc.var * kTcCols + pc + fc.chan];
            float part_ = seed * cf * dpow_i(x, fc.ipow, fc.pow);
            for (int fj = tm.fac_begin; fj < tm.fac_end; ++fj) {
              if (fj == fi) continue;
              const tdb200_factor fo = sm.facS[fj];
              part_ *= pow_i((sbase + kOffU)[fo.var * kTcCols + pc + fo.chan], fo.ipow, fo.pow);
            }
            (sbase + kOffGu)[fc.var * kTcCols + pc + fc.chan] += part_;
            full *= pow_i(x, fc.ipow, fc.pow);
          }
          if (tm.kind == 2) atomicAdd(&(sbase + kOffCg)[tm.idx], seed * full);
        }
        TMARK(13);
      }
    }
    TMARK(14);
    __syncthreads();
    if (!a.do_grad) continue;

    TMARK(6);
    // ---- backward of the last layer: dWl, dbl accumulators; gY of the last tanh layer ----------------------
    if (tid < n_out) {
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += (sbase + kOffGu)[tid * kTcCols + (p / PH) * kTcPC + (p % PH) * J];
      dbl_acc += s;
    }
    float gy[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) gy[j] = 0.f;
    for (int v = 0; v < n_out; ++v) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 g4 = *reinterpret_cast<const float4*>((sbase + kOffGu) + v * kTcCols + col0 + 4 * q);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s = fmaf(g[i], y[4 * q + i], s);
          gy[4 * q + i] = fmaf(wl[v], g[i], gy[4 * q + i]);
        }
      }
#pragma unroll
      for (int vv = 0; vv < kTcMaxOut; ++vv) if (vv == v) dwl_acc[vv] += s;
    }

    // ---- backward sweep over the tanh layers t = NMMA .. 0 ------------------------------------------
    // The weight-gradient GEMM of layer t + 1 is issued by warp 0 in 2 PH portions during the epilogue of layer t.
#pragma unroll
    for (int t = NMMA; t >= 0; --t) {
      if (t < NMMA) {
        float g2[16];
        tmem_ld16(t_lane + kTmD + (uint32_t)col0, gy);
        tmem_ld16(t_lane + kTmD + (uint32_t)(kTcCols + col0), g2);
#pragma unroll
        for (int j = 0; j < 16; ++j) gy[j] += g2[j];
      }
      const uint32_t wg_d = tmem + kTmDw + (uint32_t)t * kTmDwCols;       // dW slot of layer t + 1
      float gz[16];
      float db = 0.f;
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        if (t < NMMA && warp == 0) {
          if (p == 0) issue_wgrad_range<0, 24 * 1 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 1) issue_wgrad_range<24 * 2 / (2 * PH), 24 * 3 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 2) issue_wgrad_range<24 * 4 / (2 * PH), 24 * 5 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 3) issue_wgrad_range<24 * 6 / (2 * PH), 24 * 7 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 4) issue_wgrad_range<24 * 8 / (2 * PH), 24 * 9 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 5) issue_wgrad_range<24 * 10 / (2 * PH), 24 * 11 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 6) issue_wgrad_range<24 * 12 / (2 * PH), 24 * 13 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 7) issue_wgrad_range<24 * 14 / (2 * PH), 24 * 15 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
        }
        const TanhF f(as[t][p]);
        float g0 = gy[p * J] * f.f1;
        int c = 1;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          float zz[4] = {0.f, 0.f, 0.f, 0.f}, gg[4];
          if (t == 0) zz[0] = w0d[i];
          else {
#pragma unroll
            for (int k = 0; k < ORD[i]; ++k) zz[k] = zd[t > 0 ? t - 1 : 0][p][c - 1 + k];
          }
          g0 += tanh_jet_bwd(f, zz, gy + p * J + c, ORD[i], gg);
          if (t == 0) dw0_dir[i] += gg[0];
#pragma unroll
          for (int k = 0; k < ORD[i]; ++k) gz[p * J + c + k] = gg[k];
          c += ORD[i];
        }
        gz[p * J] = g0;
        db += g0;
        if (t == 0) {
          const float4 x4 = *reinterpret_cast<const float4*>(xcur + (part * PH + p) * 4);
          dw0_acc[0] = fmaf(g0, x4.x, dw0_acc[0]); dw0_acc[1] = fmaf(g0, x4.y, dw0_acc[1]);
          dw0_acc[2] = fmaf(g0, x4.z, dw0_acc[2]); dw0_acc[3] = fmaf(g0, x4.w, dw0_acc[3]);
        }
        if (t < NMMA && warp == 0) {
          if (p == 0) issue_wgrad_range<24 * 1 / (2 * PH), 24 * 2 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 1) issue_wgrad_range<24 * 3 / (2 * PH), 24 * 4 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 2) issue_wgrad_range<24 * 5 / (2 * PH), 24 * 6 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 3) issue_wgrad_range<24 * 7 / (2 * PH), 24 * 8 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 4) issue_wgrad_range<24 * 9 / (2 * PH), 24 * 10 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 5) issue_wgrad_range<24 * 11 / (2 * PH), 24 * 12 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 6) issue_wgrad_range<24 * 13 / (2 * PH), 24 * 14 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (p == 7) issue_wgrad_range<24 * 15 / (2 * PH), 24 * 16 / (2 * PH)>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
        }
      }
      if (t < NMMA) {
        if (warp == 0) {
          if (PH > 8) issue_wgrad_range<24 * 16 / (2 * PH), 24>(wg_d, tmem + kTmAHi, tmem + kTmALo, sbase + kOffYwHi, sbase + kOffYwLo, dw_started);
          if (elect_one()) umma_commit(sm.gbar);
          __syncwarp();
        }
        wgrad_pending = true;
      }
      db_acc[t] += db;
      if (t == 0) break;
#pragma unroll
      for (int j = C; j < 16; ++j) gz[j] = 0.f;
      // gZ -> shared memory (B operand of the backward-data GEMM) and TMEM (A operand of the weight-gradient GEMM);
      // Y_{t-1} -> shared memory as the K-major B operand of the weight-gradient GEMM.  All hi / lo tf32 pairs.
      float yp[16];
      jets_from_saved(as[t - 1], zd[t > 1 ? t - 2 : 0], t == 1, yp);
      if (wgrad_pending) { mbar_wait(sm.gbar, gphase); gphase ^= 1; wgrad_pending = false; tc_fence_after(); }
      {
        float hi[16], lo[16];
        if (!live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) gz[j] = 0.f;
        }
        split16(gz, hi, lo);
        if (live) {
          st4((sbase + kOffActHi) + actA, hi); st4((sbase + kOffActHi) + actA + 4, hi + 4); st4((sbase + kOffActHi) + actB, hi + 8); st4((sbase + kOffActHi) + actB + 4, hi + 12);
          st4((sbase + kOffActLo) + actA, lo); st4((sbase + kOffActLo) + actA + 4, lo + 4); st4((sbase + kOffActLo) + actB, lo + 8); st4((sbase + kOffActLo) + actB + 4, lo + 12);
        }
        tmem_st16(t_lane + kTmAHi + (uint32_t)col0, hi);
        tmem_st16(t_lane + kTmALo + (uint32_t)col0, lo);
        if (live) {
          split16(yp, hi, lo);
#pragma unroll
          for (int q = 0; q < 4; ++q) { st4((sbase + kOffYwHi) + ywq[q], hi + 4 * q); st4((sbase + kOffYwLo) + ywq[q], lo + 4 * q); }
        }
        tmem_st_wait();
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      TMARK(7);
      if (warp == 0) {
        mbar_wait(sm.wbar, wphase); wphase ^= 1;        // W_t^T image has landed
        tc_fence_after();
        issue_gemm_any(tmem + kTmD, (sbase + kOffWHi), (sbase + kOffWLo), (sbase + kOffActHi), (sbase + kOffActLo), ksteps);   // A = W_t^T image
        if (elect_one()) umma_commit(sm.bar);
        __syncwarp();
        TMARK(8);
      }
      mbar_wait(sm.bar, phase);
      phase ^= 1;
      tc_fence_after();
      TMARK(9);
      if (tid == 0) {                                   // next image: W_{t-1}^T, or W_1 for the next tile
        const float* nxt = t > 1 ? wimg + (size_t)(t - 2) * 4 * kTcWFloats + 2 * kTcWFloats : wimg;
        bulk_load_image((sbase + kOffWHi), nxt, sm.wbar);
      }
    }
    dw_started = 1;
    TMARK(10);
  }
#ifdef TDB_TC_TIMING
  if (a.dbg && tid == 0)
    for (int i = 0; i < 16; ++i) a.dbg[(size_t)blockIdx.x * 16 + i] = tacc[i];
#endif

  // ---- flush: per-thread accumulators, dW accumulators (TMEM), per-CTA scalars --------------------------
  if (warp == 0) { mbar_wait(sm.wbar, wphase); wphase ^= 1; }   // drain the last prefetch before exiting
  if (wgrad_pending) { mbar_wait(sm.gbar, gphase); gphase ^= 1; wgrad_pending = false; }
  __syncthreads();
  tc_fence_after();
  if (a.do_grad) {
    // per-thread accumulators of the four column parts -> staging in the (now free) weight buffer -> part 0 adds them up
    float* const stg = sbase + kOffWHi;                   // [4 parts][NMMA + 1 + 4 + kTcMaxOut][128]
    constexpr int kRowsStg = NMMA + 1 + 4 + kTcMaxOut;
    for (int i = 0; i < ND; ++i)                         // derivative-channel part of dW0, by jet direction
      for (int ax = 0; ax < 4; ++ax) dw0_acc[ax] = fmaf(dw0_dir[i], sg.dir_vec[i][ax], dw0_acc[ax]);
    {
      float* mine = stg + (part * kRowsStg) * 128 + n;
#pragma unroll
      for (int l = 0; l <= NMMA; ++l) mine[l * 128] = db_acc[l];
#pragma unroll
      for (int ax = 0; ax < 4; ++ax) mine[(NMMA + 1 + ax) * 128] = dw0_acc[ax];
#pragma unroll
      for (int v = 0; v < kTcMaxOut; ++v) mine[(NMMA + 5 + v) * 128] = dwl_acc[v];
    }
    __syncthreads();
    if (live && part == 0) {
      auto total = [&](int r) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < kTcParts; ++q) t += stg[(q * kRowsStg + r) * 128 + n];
        return t;
      };
#pragma unroll
      for (int l = 0; l <= NMMA; ++l) my_grad[a.b_off[l] + n] = total(l);
      for (int ax = 0; ax < d; ++ax) my_grad[a.w_off[0] + n * d + ax] = total(NMMA + 1 + ax);
#pragma unroll
      for (int v = 0; v < kTcMaxOut; ++v) if (v < n_out) my_grad[a.w_off[L - 1] + v * W + n] = total(NMMA + 5 + v);
    }
    if (tid < n_out) my_grad[a.b_off[L - 1] + tid] = dbl_acc;
    if (dw_started) {
      float* row0 = my_grad;
      for (int t = 1; t <= NMMA; ++t) {
        float* dst = row0 + a.w_off[t];
        for (int k0 = part * 32; k0 < part * 32 + 32 && k0 < (int)kTmDwCols; k0 += 16) {
          float v[16];
          tmem_ld16(t_lane + kTmDw + (uint32_t)(t - 1) * kTmDwCols + (uint32_t)k0, v);
          if (live)
            for (int j = 0; j < 16; ++j)
              if (k0 + j < W) dst[(size_t)n * W + k0 + j] = v[j];
        }
      }
    }
  }
  if (tid < a.n_slots) {
    double s = 0.0;
    for (int col = 0; col < ncols; ++col)
      if (sg.col_slot[col] == tid)
        for (int p = 0; p < P; ++p) s += sm.lossT[p * TDB200_MAX_COLS + col];
    a.part_loss[(size_t)blockIdx.x * a.n_slots + tid] = s;
  }
  if (a.do_grad && tid < a.n_cparams) my_grad[a.n_net_params + tid] = (sbase + kOffCg)[tid];   // warp 0 -> part-0 row
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------
// launch of one signature; the instantiations are spread over jet_tc.cu / jet_tc_g1..3.cu (parallel compilation)
// ------------------------------------------------------------------------------------------------
template <int O0, int O1, int O2, int NMMA>
static cudaError_t launch_sig(const JetArgs& a, const float* wimg, int grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(jet_tc_kernel<O0, O1, O2, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kTcSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  jet_tc_kernel<O0, O1, O2, NMMA><<<grid, kTcThreads, kTcSmemBytes, s>>>(a, wimg);
  return cudaGetLastError();
}

#define TDB_TC_SIGS_G0(X) X(0, 0, 0) X(1, 0, 0) X(2, 0, 0) X(3, 0, 0) X(4, 0, 0) X(1, 1, 0)
#define TDB_TC_SIGS_G1(X) X(2, 1, 0) X(1, 2, 0) X(2, 2, 0) X(3, 1, 0) X(1, 3, 0)
#define TDB_TC_SIGS_G2(X) X(3, 2, 0) X(2, 3, 0) X(4, 1, 0) X(1, 4, 0) X(4, 2, 0) X(2, 4, 0)
#define TDB_TC_SIGS_G3(X) X(3, 3, 0) X(1, 1, 1) X(2, 1, 1) X(2, 2, 1) X(2, 2, 2)
#define TDB_TC_SIGS(X) TDB_TC_SIGS_G0(X) TDB_TC_SIGS_G1(X) TDB_TC_SIGS_G2(X) TDB_TC_SIGS_G3(X)

// one group of signatures per translation unit: returns cudaErrorInvalidValue when the signature is not in the group
#define TDB_TC_DEFINE_GROUP(NAME, SIGS)                                                                         \
  cudaError_t NAME(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s) {     \
    const int n_mma = a.n_layers - 2;                                                                           \
    SIGS(TDB_TC_GROUP_CASE)                                                                                     \
    return cudaErrorInvalidValue;                                                                               \
  }
#define TDB_TC_GROUP_CASE(A, B, Cc)                                                                             \
  if (o0 == A && o1 == B && o2 == Cc)                                                                           \
    return n_mma == 1 ? launch_sig<A, B, Cc, 1>(a, wimg, grid, s) : launch_sig<A, B, Cc, 2>(a, wimg, grid, s);

cudaError_t launch_jet_tc_g0(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_jet_tc_g1(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_jet_tc_g2(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_jet_tc_g3(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s);

}  // namespace tdb
