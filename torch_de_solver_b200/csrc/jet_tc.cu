// Fused Taylor-jet MLP loss + gradient kernel on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as jet_simt.cu for the interior (identity, K = 1) segments of nets whose hidden layers all have
// the same width W <= 104; the W x W layer GEMMs (forward, backward-data, weight gradient) run as
// tcgen05.mma.kind::tf32 with fp32 accumulators in TMEM, in the 3xTF32 split (hi*hi + hi*lo + lo*hi).
//
// Orientation: D[neuron, (point, channel)] = W . Y^T, i.e. accumulator lanes are neurons and columns are the
// (point, jet-channel) pairs of the tile.  A thread that owns lane n therefore sees every jet channel of every
// point for its neuron, so the tanh-jet rule and its adjoint are thread-local TMEM -> registers -> smem epilogues.
//   forward      D[n,(pc)]  = sum_k W[n,k]   Y[(pc),k]     A = W image   (smem, K-major), B = activations (K-major)
//   backward     D[k,(pc)]  = sum_n W^T[k,n] gZ[(pc),n]    A = W^T image (smem, K-major), B = gZ (K-major)
//   weight grad dW[n,k]    += sum_pc gZ[n,(pc)] Y[(pc),k]   A = gZ straight from the epilogue registers into TMEM
//                                                          (tcgen05.st), B = Y (smem, MN-major); dW stays in TMEM
//                                                          for the whole kernel and is flushed once per CTA.
// K-major operands use the canonical 128-byte swizzle ([k-block of 32][row][32 floats], 16-byte chunks XOR row);
// the MN-major tf32 operand needs the 32-byte-base variant (SWIZZLE_128B_BASE32B: 4-row atoms, 32-byte chunks).
#include "common.cuh"

namespace tdb {

constexpr int kTcThreads = 512;                  // 16 warps: 4 lane windows x 4 column parts
constexpr int kTcParts = 4;
constexpr int kTcCols = 48;                      // (point, channel) columns per tile = MMA N
constexpr int kTcWRows = 104;                    // rows of the weight image (neurons padded to 8)
constexpr int kTcActBlock = kTcCols * 32;        // floats per k-block of an activation operand
constexpr int kTcActFloats = 4 * kTcActBlock;    // 6144 floats = 24 KB
constexpr int kTcWBlock = kTcWRows * 32;
constexpr int kTcWFloats = 4 * kTcWBlock;        // 13312 floats = 52 KB
constexpr int kTcMaxMma = 2;                     // W x W layers: Z_l, dW_l and the operands all stay in TMEM
constexpr int kTcSavePitch = 104;
// TMEM columns: Z_1 | Z_2 (forward accumulators, kept for the backward sweep) | D_bwd | gZ hi | gZ lo (A operands
// of the weight-gradient MMA) | dW slots (112 columns each)
constexpr uint32_t kTmZ = 0, kTmDb = 96, kTmAHi = 144, kTmALo = 192, kTmDw = 240, kTmDwCols = 112;

// float offset of element (row, k) inside a swizzled operand buffer with `rows` rows per k-block
__host__ __device__ __forceinline__ int sw_off(int row, int k, int rows) {
  return (k >> 5) * rows * 32 + row * 32 + ((((k & 31) >> 2) ^ (row & 7)) << 2) + (k & 3);
}

// float offset of element (K-row r, MN index k) of an MN-major tf32 operand (SWIZZLE_128B_BASE32B)
__host__ __device__ __forceinline__ int sw_off_mn(int r, int k, int rows) {
  return (k >> 5) * rows * 32 + r * 32 + ((((k & 31) >> 3) ^ (r & 3)) << 3) + (k & 7);
}

// ------------------------------------------------------------------------------------------------
// weight images: [layer][W hi | W lo | W^T hi | W^T lo][4 k-blocks][104 rows][32]; hi = tf32(w) (rna), lo = w - hi
// ------------------------------------------------------------------------------------------------
__global__ void pack_tc_images_kernel(PackArgs a, float* __restrict__ img) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  for (int l = 1; l <= a.n_layers - 2; ++l) {
    const int in = a.widths[l], out = a.widths[l + 1];
    float* hi = img + (size_t)(l - 1) * 4 * kTcWFloats;
    float* lo = hi + kTcWFloats;
    float* thi = lo + kTcWFloats;
    float* tlo = thi + kTcWFloats;
    const float* __restrict__ W = a.W[l];
    for (int i = tid; i < in * out; i += nth) {
      const int n = i / in, k = i - n * in;
      const float w = W[i];
      uint32_t hb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(w));
      const float h = __uint_as_float(hb);
      const int o = sw_off(n, k, kTcWRows), ot = sw_off(k, n, kTcWRows);
      hi[o] = h;
      lo[o] = w - h;
      thi[ot] = h;
      tlo[ot] = w - h;
    }
  }
}

cudaError_t launch_pack_tc_images(const PackArgs& a, float* img, cudaStream_t s) {
  pack_tc_images_kernel<<<32, 256, 0, s>>>(a, img);
  return cudaGetLastError();
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
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

// ------------------------------------------------------------------------------------------------
// MMA issue helpers (one thread).  All operand buffers are 1024-byte aligned.
// ------------------------------------------------------------------------------------------------
// The issue helpers are called by every lane of one warp (warp-uniform control flow keeps the descriptors in
// uniform registers); each tcgen05.mma itself is issued by the elected lane.
// D[128 x 48] (+)= A(W image, K-major) . B(act, K-major), 3xTF32.  KS = K-steps of 8 (compile time).
template <int KS>
__device__ __forceinline__ void issue_forward(uint32_t d_tmem, const float* w_hi, const float* w_lo,
                                              const float* b_hi, const float* b_lo) {
  constexpr uint32_t idesc = umma_idesc(128, kTcCols, 0, 0);
  const uint64_t dwh = umma_desc(smem_u32(w_hi), 16, 1024), dwl = umma_desc(smem_u32(w_lo), 16, 1024);
  const uint64_t dbh = umma_desc(smem_u32(b_hi), 16, 1024), dbl = umma_desc(smem_u32(b_lo), 16, 1024);
  const bool leader = elect_one();
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint64_t A = pass == 0 ? dwl : dwh;        // lo*hi, hi*lo, hi*hi
    const uint64_t B = pass == 1 ? dbl : dbh;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const uint64_t ao = ((uint64_t)(s >> 2) * kTcWBlock * 4 + (uint64_t)(s & 3) * 32) >> 4;
      const uint64_t bo = ((uint64_t)(s >> 2) * kTcActBlock * 4 + (uint64_t)(s & 3) * 32) >> 4;
      if (leader) umma_tf32(d_tmem, A + ao, B + bo, idesc, (pass | s) ? 1u : 0u);
    }
  }
}
__device__ __forceinline__ void issue_forward_any(uint32_t d_tmem, const float* w_hi, const float* w_lo,
                                                  const float* b_hi, const float* b_lo, int ksteps) {
  if (ksteps == 13) issue_forward<13>(d_tmem, w_hi, w_lo, b_hi, b_lo);
  else if (ksteps <= 4) issue_forward<4>(d_tmem, w_hi, w_lo, b_hi, b_lo);     // zero-padded images: extra steps add 0
  else if (ksteps <= 8) issue_forward<8>(d_tmem, w_hi, w_lo, b_hi, b_lo);
  else issue_forward<13>(d_tmem, w_hi, w_lo, b_hi, b_lo);
}
// dW[128 (n) x 112 (k)] += A(gZ in TMEM: lanes n, columns (pc)) . B(Y in smem read MN-major: N = k, K = (pc))
__device__ __forceinline__ void issue_wgrad(uint32_t d_tmem, uint32_t a_hi_tmem, uint32_t a_lo_tmem,
                                            const float* y_hi, const float* y_lo, uint32_t accumulate) {
  constexpr uint32_t idesc = umma_idesc(128, 112, 0, 1);
  // 8 (pc) rows = two 4-row swizzle atoms (SBO = 512 B); MN blocks of 32 k at LBO = one k-block
  const uint64_t dyh = umma_desc(smem_u32(y_hi), kTcActBlock * 4, 512, 1), dyl = umma_desc(smem_u32(y_lo), kTcActBlock * 4, 512, 1);
  const bool leader = elect_one();
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t A = pass == 0 ? a_lo_tmem : a_hi_tmem;
    const uint64_t B = pass == 1 ? dyl : dyh;
#pragma unroll
    for (int s = 0; s < kTcCols / 8; ++s)
      if (leader) umma_tf32_ts(d_tmem, A + (uint32_t)s * 8, B + (uint64_t)(s * 1024 >> 4), idesc, (pass | s) ? 1u : accumulate);
  }
}

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
// tanh with ~1e-7 absolute error: odd polynomial near 0 (Cephes tanhf), 1 - 2 / (exp(2|x|) + 1) elsewhere
__device__ __forceinline__ float tanh_acc(float x) {
  const float ax = fabsf(x);
  if (ax < 0.625f) {
    const float s = x * x;
    const float p = ((((-5.70498872745e-3f * s + 2.06390887954e-2f) * s - 5.37397155531e-2f) * s +
                      1.33314422036e-1f) * s - 3.33332819422e-1f);
    return fmaf(x * s, p, x);
  }
  const float e = __expf(2.f * ax);
  return copysignf(1.f - __fdividef(2.f, e + 1.f), x);
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 consecutive columns of this warp's lane window (a part owns at most 12 of them)
__device__ __forceinline__ void tmem_ld16w(uint32_t taddr, float* v) {
  tmem_ld8_nowait(taddr, v);
  tmem_ld8_nowait(taddr + 8, v + 8);
  tmem_ld_wait();
}
// exactly C (compile time, <= 24) columns
template <int C>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const float* v) {
  int c = 0;
#pragma unroll
  for (; c + 4 <= C; c += 4) tmem_st4(taddr + c, v + c);
  if (C & 2) { tmem_st2(taddr + c, v + c); c += 2; }
  if (C & 1) tmem_st1(taddr + c, v[c]);
}

// warp-wide sum of 32 per-lane values: afterwards lane L holds the total of v[L] over the warp (31 shuffles)
__device__ __forceinline__ float warp_multi_reduce32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = up ? v[i + off] : v[i];
      const float send = up ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
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
struct TcSmem {
  float *w_hi, *w_lo, *a_hi, *a_lo, *b_hi, *b_lo;
  float *xS, *uS, *guS, *uP, *wlS, *cgS;
  tdb200_term* termS;
  tdb200_factor* facS;
  tdb200_segment* segS;
  float* scaleS;          // [32] lambda / len per slot
  double* lossT;          // [kTcCols][TDB200_MAX_COLS] per-point-thread loss accumulators (no atomics)
  double* lossS;
  uint64_t *bar, *wbar, *gbar;
  uint32_t* tmem_ptr;
};
constexpr int kTcMaxTerms = 48, kTcMaxFactors = 96;   // operator program cached in shared memory
constexpr size_t kTcSmemBytes = (size_t)(2 * kTcWFloats + 4 * kTcActFloats) * 4 + 1024 /*align*/ +
                                (2 * kTcCols * 4 + 2 * kMaxOut * kTcCols + 4 * kMaxOut * kTcCols +
                                 kMaxOut * kTcSavePitch + kMaxCParams) * 4 + 32 * 8 + 64 +
                                kTcMaxTerms * sizeof(tdb200_term) + kTcMaxFactors * sizeof(tdb200_factor) + 16 +
                                sizeof(tdb200_segment) + 32 * 4 + kTcCols * TDB200_MAX_COLS * 8 + 32;

size_t jet_tc_smem_bytes() { return kTcSmemBytes; }

// Jet signature: derivative orders of up to three directions (sorted by input axis); J = 1 + O0 + O1 + O2.
template <int O0, int O1, int O2, int NMMA>
__global__ void __launch_bounds__(kTcThreads, 1) jet_tc_kernel(const JetArgs a, const float* __restrict__ wimg) {
  constexpr int J = 1 + O0 + O1 + O2;
  constexpr int ND = (O0 > 0) + (O1 > 0) + (O2 > 0);
  constexpr int PH = 12 / J;                   // points per column part
  constexpr int P = kTcParts * PH;             // points per tile
  constexpr int C = PH * J;                    // columns per part (<= 12)
  constexpr int n_mma = NMMA;                  // W x W layers (1 or 2)
  constexpr int ORD[3] = {O0, O1, O2};
  extern __shared__ uint8_t smem_raw_tc[];
  TcSmem sm;
  {
    uintptr_t base = (reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~uintptr_t(1023);
    float* f = reinterpret_cast<float*>(base);
    sm.w_hi = f; f += kTcWFloats;
    sm.w_lo = f; f += kTcWFloats;
    sm.a_hi = f; f += kTcActFloats;
    sm.a_lo = f; f += kTcActFloats;
    sm.b_hi = f; f += kTcActFloats;
    sm.b_lo = f; f += kTcActFloats;
    sm.xS = f; f += 2 * kTcCols * 4;                   // double buffered: the next tile's points are prefetched
    sm.uS = f; f += kMaxOut * kTcCols;
    sm.guS = f; f += kMaxOut * kTcCols;
    sm.uP = f; f += 4 * kMaxOut * kTcCols;
    sm.wlS = f; f += kMaxOut * kTcSavePitch;
    sm.cgS = f; f += kMaxCParams;
    sm.lossS = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(f) + 15) & ~uintptr_t(15));
    sm.bar = reinterpret_cast<uint64_t*>(sm.lossS + 32);
    sm.wbar = sm.bar + 1;
    sm.gbar = sm.bar + 2;
    sm.tmem_ptr = reinterpret_cast<uint32_t*>(sm.bar + 3);
    sm.termS = reinterpret_cast<tdb200_term*>((reinterpret_cast<uintptr_t>(sm.tmem_ptr + 2) + 15) & ~uintptr_t(15));
    sm.facS = reinterpret_cast<tdb200_factor*>(sm.termS + kTcMaxTerms);
    sm.segS = reinterpret_cast<tdb200_segment*>(sm.facS + kTcMaxFactors);
    sm.scaleS = reinterpret_cast<float*>(sm.segS + 1);
    sm.lossT = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm.scaleS + 32) + 15) & ~uintptr_t(15));
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = (warp & 3) * 32 + lane;                 // neuron = TMEM lane owned by this thread
  const int half = warp >> 2;                           // which column part (0..3) of the tile this thread owns
  const int L = a.n_layers, W = a.widths[1], n_out = a.widths[L], d = a.d;
  const int ksteps = (W + 7) / 8;
  const bool live = n < W;
  const int col0 = half * C;                            // first (point, channel) column of this thread
  // one gradient-partial row per column part: a single owner thread per address -> bit-reproducible
  float* const my_grad = a.part_grad + ((size_t)blockIdx.x * kTcParts + half) * a.n_params_pad;

  // ---- one-time setup --------------------------------------------------------------------------------
  for (int i = tid; i < kTcParts * a.n_params_pad; i += kTcThreads)
    a.part_grad[(size_t)blockIdx.x * kTcParts * a.n_params_pad + i] = 0.f;
  for (int i = tid; i < 4 * kTcActFloats; i += kTcThreads) sm.a_hi[i] = 0.f;      // pad rows / columns stay zero
  if (tid < 32) sm.lossS[tid] = 0.0;
  if (tid < kMaxCParams) sm.cgS[tid] = 0.f;
  for (int i = tid; i < n_out * W; i += kTcThreads) sm.wlS[(i / W) * kTcSavePitch + i % W] = a.arena[a.w_off[L - 1] + i];
  for (int i = tid; i < min(kTcMaxTerms, a.n_terms); i += kTcThreads) sm.termS[i] = a.terms[i];
  for (int i = tid; i < min(kTcMaxFactors, a.n_factors); i += kTcThreads) sm.facS[i] = a.factors[i];
  for (int i = tid; i < (int)(sizeof(tdb200_segment) / 4); i += kTcThreads)
    reinterpret_cast<uint32_t*>(sm.segS)[i] = reinterpret_cast<const uint32_t*>(a.segs)[i];
  if (tid < a.n_slots) sm.scaleS[tid] = a.slot_scale[tid];
  for (int i = tid; i < kTcCols * TDB200_MAX_COLS; i += kTcThreads) sm.lossT[i] = 0.0;
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
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);    // this warp's lane window
  {  // the gZ operand columns of TMEM must hold zeros where no (point, channel) column exists
    const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (half == 0)
      for (uint32_t c = kTmAHi; c < kTmDw; c += 8) { tmem_st4(t_lane + c, z8); tmem_st4(t_lane + c + 4, z8); }
    tmem_st_wait();
  }
  long long tacc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tacc[i] = 0;
  long long tlast = clock64();
#define TMARK(i) do { if (a.dbg) { const long long tn_ = clock64(); tacc[i] += tn_ - tlast; tlast = tn_; } } while (0)
  uint32_t phase = 0, wphase = 0, gphase = 0;           // wphase is only used by warp 0
  bool wgrad_pending = false;                           // weight-gradient MMAs still reading TMEM A / Y operand
  uint32_t dw_started = 0;
  // per-layer parameters this thread needs all the time
  float bias[3] = {0.f, 0.f, 0.f}, w0[4] = {0.f, 0.f, 0.f, 0.f}, wl[kMaxOut];
  if (live) {
    for (int l = 0; l <= n_mma; ++l) bias[l] = a.arena[a.b_off[l] + n];
    for (int ax = 0; ax < d; ++ax) w0[ax] = a.arena[a.w_off[0] + n * d + ax];
  }
#pragma unroll
  for (int v = 0; v < kMaxOut; ++v) wl[v] = (live && v < n_out) ? a.arena[a.w_off[L - 1] + v * W + n] : 0.f;

  const tdb200_segment& sg = *sm.segS;                  // shared-memory copy (set up above, visible after the sync)
  const int ncols = sg.n_cols;
  int dir_axis[3] = {0, 0, 0};
  for (int i = 0; i < ND; ++i) dir_axis[i] = sg.dir_axis[i];

  if (tid == 0) bulk_load_image(sm.w_hi, wimg, sm.wbar);          // W_1 for the first tile

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
  load_points(blockIdx.x, sm.xS);
  int xbuf = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const long long g_first = (long long)tile * P;
    const int p_valid = (int)min((long long)P, sg.n_groups - g_first);
    float* const xcur = sm.xS + xbuf * kTcCols * 4;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                    // this tile's points (loaded one tile ahead) are visible
    if (tile + (int)gridDim.x < a.n_tiles) load_points(tile + gridDim.x, sm.xS + (xbuf ^ 1) * kTcCols * 4);
    xbuf ^= 1;
    TMARK(0);

    float yk[NMMA + 1][12];                             // outputs of tanh layers 0..n_mma for this thread's columns

    // ---- layer 0 (K = d): thread-local ----------------------------------------------------------------
#pragma unroll
    for (int p = 0; p < PH; ++p) {
      float z0 = bias[0];
      for (int ax = 0; ax < d; ++ax) z0 = fmaf(w0[ax], xcur[(half * PH + p) * 4 + ax], z0);
      const float av = tanh_acc(z0);
      const TanhF f(av);
      yk[0][p * J] = av;
      int c = 1;
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        float z[4] = {w0[dir_axis[i]], 0.f, 0.f, 0.f}, y[4];
        tanh_jet_fwd(f, z, ORD[i], y);
#pragma unroll
        for (int k = 0; k < ORD[i]; ++k) yk[0][p * J + c + k] = y[k];
        c += ORD[i];
      }
    }
    if (wgrad_pending) { mbar_wait(sm.gbar, gphase); gphase ^= 1; wgrad_pending = false; }   // Y operand is free again
    if (live) {
#pragma unroll
      for (int j = 0; j < C; ++j) split_store(sm.a_hi, sm.a_lo, sw_off(col0 + j, n, kTcCols), yk[0][j]);
    }

    // ---- W x W layers: tensor-core GEMM + thread-local tanh-jet epilogue ------------------------------
#pragma unroll
    for (int l = 1; l <= n_mma; ++l) {
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      TMARK(1);
      if (warp == 0) {
        mbar_wait(sm.wbar, wphase); wphase ^= 1;        // W_l image has landed
        TMARK(2);
        tc_fence_after();
        issue_forward_any(tmem + kTmZ + 48u * (uint32_t)(l - 1), sm.w_hi, sm.w_lo, sm.a_hi, sm.a_lo, ksteps);
        if (elect_one()) umma_commit(sm.bar);
        __syncwarp();
        TMARK(3);
      }
      mbar_wait(sm.bar, phase);
      phase ^= 1;
      tc_fence_after();
      TMARK(4);
      // next image (W_{l+1}, or W_{n_mma}^T for the backward sweep) streams in behind the epilogue
      if (tid == 0)
        bulk_load_image(sm.w_hi, wimg + (size_t)(l < n_mma ? l : n_mma - 1) * 4 * kTcWFloats + (l < n_mma ? 0 : 2 * kTcWFloats), sm.wbar);
      float z[16];
      tmem_ld16w(t_lane + kTmZ + 48u * (uint32_t)(l - 1) + (uint32_t)col0, z);
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        const float av = tanh_acc(z[p * J] + bias[l]);
        const TanhF f(av);
        yk[l][p * J] = av;
        int c = 1;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          float y[4];
          tanh_jet_fwd(f, z + p * J + c, ORD[i], y);
#pragma unroll
          for (int k = 0; k < ORD[i]; ++k) yk[l][p * J + c + k] = y[k];
          c += ORD[i];
        }
      }
      if (live && l < n_mma) {
#pragma unroll
        for (int j = 0; j < C; ++j) split_store(sm.a_hi, sm.a_lo, sw_off(col0 + j, n, kTcCols), yk[l][j]);
      }
    }

    TMARK(5);
    // ---- last layer: u[v][col] = sum_n Wl[v][n] y[n][col]  (warp multi-value reduction, fixed order) ------
    for (int v = 0; v < n_out; ++v) {
      float t32[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) t32[j] = (j < C && live) ? wl[v] * yk[n_mma][j] : 0.f;
      const float tot = warp_multi_reduce32(t32, lane);
      if (lane < C) sm.uP[((warp & 3) * kMaxOut + v) * kTcCols + col0 + lane] = tot;
    }
    __syncthreads();
    for (int idx = tid; idx < n_out * kTcParts * C; idx += kTcThreads) {
      const int v = idx / (kTcParts * C), r = idx - v * (kTcParts * C);
      float s = (r % J) == 0 ? a.arena[a.b_off[L - 1] + v] : 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) s += sm.uP[(w * kMaxOut + v) * kTcCols + r];
      sm.uS[v * kTcCols + r] = s;
      sm.guS[v * kTcCols + r] = 0.f;
    }
    __syncthreads();

    TMARK(6);
    // ---- operator terms, residual, loss, adjoint seeds (one thread per point) -------------------------
    if (tid < p_valid) {
      const int p = tid;
      const long long row = g_first + p;
      for (int col = 0; col < ncols; ++col) {
        float val = 0.f;
        for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
          const tdb200_term tm = sm.termS[t];
          float prod = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                               : a.arena[a.n_net_params + tm.idx];
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = sm.facS[fi];
            prod *= pow_i(sm.uS[fc.var * kTcCols + p * J + fc.chan], fc.ipow, fc.pow);
          }
          val += prod;
        }
        if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
        const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
        const float res = val - tgt;
        const int slot = sg.col_slot[col];
        sm.lossT[p * TDB200_MAX_COLS + col] += (double)res * (double)res;
        if (!a.do_grad) continue;
        const float seed = 2.f * sm.scaleS[slot] * res;
        for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
          const tdb200_term tm = sm.termS[t];
          const float cf = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                                   : a.arena[a.n_net_params + tm.idx];
          float full = 1.f;
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = sm.facS[fi];
            const float x = sm.uS[fc.var * kTcCols + p * J + fc.chan];
            float part = seed * cf * dpow_i(x, fc.ipow, fc.pow);
            for (int fj = tm.fac_begin; fj < tm.fac_end; ++fj) {
              if (fj == fi) continue;
              const tdb200_factor fo = sm.facS[fj];
              part *= pow_i(sm.uS[fo.var * kTcCols + p * J + fo.chan], fo.ipow, fo.pow);
            }
            sm.guS[fc.var * kTcCols + p * J + fc.chan] += part;
            full *= pow_i(x, fc.ipow, fc.pow);
          }
          if (tm.kind == 2) atomicAdd(&sm.cgS[tm.idx], seed * full);
        }
      }
    }
    __syncthreads();
    if (!a.do_grad) {
      // forward-only evaluation: the image prefetched for the backward sweep is not needed; fetch W_1 instead
      if (warp == 0) { mbar_wait(sm.wbar, wphase); wphase ^= 1; __syncwarp(); if (tid == 0) bulk_load_image(sm.w_hi, wimg, sm.wbar); }
      continue;
    }

    TMARK(7);
    // ---- backward of the last layer: dWl, dbl; gY of the last tanh layer ---------------------------------
    if (tid < n_out) {                                   // warp 0 -> part-0 row
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += sm.guS[tid * kTcCols + p * J];
      atomicAdd(my_grad + a.b_off[L - 1] + tid, s);
    }
    float gy[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) gy[j] = 0.f;
    for (int v = 0; v < n_out; ++v) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float g = sm.guS[v * kTcCols + col0 + j];
        s = fmaf(g, yk[n_mma][j], s);
        gy[j] = fmaf(wl[v], g, gy[j]);
      }
      if (live) atomicAdd(my_grad + a.w_off[L - 1] + v * W + n, s);
    }

    // ---- backward sweep over the tanh layers t = n_mma .. 0 ------------------------------------------
#pragma unroll
    for (int t = n_mma; t >= 0; --t) {
      float z[16];
      if (t > 0) tmem_ld16w(t_lane + kTmZ + 48u * (uint32_t)(t - 1) + (uint32_t)col0, z);
      if (t < n_mma) tmem_ld16w(t_lane + kTmDb + (uint32_t)col0, gy);
      float gz[12];
      float db = 0.f, dw0[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        const float av = yk[t][p * J];
        const TanhF f(av);
        float g0 = gy[p * J] * f.f1;
        int c = 1;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          float zz[4] = {0.f, 0.f, 0.f, 0.f}, gg[4];
          if (t == 0) zz[0] = w0[dir_axis[i]];
          else {
#pragma unroll
            for (int k = 0; k < ORD[i]; ++k) zz[k] = z[p * J + c + k];
          }
          g0 += tanh_jet_bwd(f, zz, gy + p * J + c, ORD[i], gg);
          if (t == 0) dw0[dir_axis[i]] += gg[0];
#pragma unroll
          for (int k = 0; k < ORD[i]; ++k) gz[p * J + c + k] = gg[k];
          c += ORD[i];
        }
        gz[p * J] = g0;
        db += g0;
        if (t == 0)
          for (int ax = 0; ax < d; ++ax) dw0[ax] = fmaf(g0, xcur[(half * PH + p) * 4 + ax], dw0[ax]);
      }
      if (live) atomicAdd(my_grad + a.b_off[t] + n, db);
      if (t == 0) {
        if (live) for (int ax = 0; ax < d; ++ax) atomicAdd(my_grad + a.w_off[0] + n * d + ax, dw0[ax]);
        break;
      }
      // gZ -> shared memory (B operand of the backward-data GEMM) and TMEM (A operand of the weight-gradient GEMM);
      // Y_{t-1} -> shared memory as the MN-major B operand of the weight-gradient GEMM.  All hi / lo tf32 pairs.
      if (wgrad_pending) { mbar_wait(sm.gbar, gphase); gphase ^= 1; wgrad_pending = false; tc_fence_after(); }
      {
        float ghi[12], glo[12];
#pragma unroll
        for (int j = 0; j < C; ++j) {
          const float g = live ? gz[j] : 0.f;
          uint32_t hb;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(g));
          ghi[j] = __uint_as_float(hb);
          glo[j] = g - ghi[j];
        }
        if (live) {
#pragma unroll
          for (int j = 0; j < C; ++j) {
            const int o = sw_off(col0 + j, n, kTcCols);
            sm.b_hi[o] = ghi[j];
            sm.b_lo[o] = glo[j];
            split_store(sm.a_hi, sm.a_lo, sw_off_mn(col0 + j, n, kTcCols), yk[t - 1][j]);
          }
        }
        tmem_st_cols<C>(t_lane + kTmAHi + (uint32_t)col0, ghi);
        tmem_st_cols<C>(t_lane + kTmALo + (uint32_t)col0, glo);
        tmem_st_wait();
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      TMARK(8);
      if (warp == 0) {
        mbar_wait(sm.wbar, wphase); wphase ^= 1;        // W_t^T image has landed
        TMARK(9);
        tc_fence_after();
        issue_forward_any(tmem + kTmDb, sm.w_hi, sm.w_lo, sm.b_hi, sm.b_lo, ksteps);   // A = W_t^T image
        if (elect_one()) umma_commit(sm.bar);
        __syncwarp();
        // the weight gradient is not on the critical path: it runs behind the next adjoint epilogue
        issue_wgrad(tmem + kTmDw + (uint32_t)(t - 1) * kTmDwCols, tmem + kTmAHi, tmem + kTmALo, sm.a_hi, sm.a_lo,
                    dw_started);
        if (elect_one()) umma_commit(sm.gbar);
        __syncwarp();
        TMARK(10);
      }
      wgrad_pending = true;
      mbar_wait(sm.bar, phase);
      phase ^= 1;
      tc_fence_after();
      TMARK(11);
      if (tid == 0) {                                   // next image: W_{t-1}^T, or W_1 for the next tile
        const float* nxt = t > 1 ? wimg + (size_t)(t - 2) * 4 * kTcWFloats + 2 * kTcWFloats : wimg;
        bulk_load_image(sm.w_hi, nxt, sm.wbar);
      }
    }
    dw_started = 1;
    tc_fence_before();
    __syncthreads();
    TMARK(12);
  }
  if (a.dbg && tid == 0)
    for (int i = 0; i < 16; ++i) a.dbg[(size_t)blockIdx.x * 16 + i] = tacc[i];

  // ---- flush: dW accumulators (TMEM) and per-CTA scalars ----------------------------------------------
  if (warp == 0) { mbar_wait(sm.wbar, wphase); wphase ^= 1; }   // drain the last prefetch before exiting
  if (wgrad_pending) { mbar_wait(sm.gbar, gphase); gphase ^= 1; wgrad_pending = false; }
  __syncthreads();
  tc_fence_after();
  if (a.do_grad && dw_started) {
    float* row0 = a.part_grad + (size_t)blockIdx.x * kTcParts * a.n_params_pad;   // dW lives in the part-0 row
    for (int t = 1; t <= n_mma; ++t) {
      float* dst = row0 + a.w_off[t];
      for (int k0 = half * 32; k0 < half * 32 + 32 && k0 < (int)kTmDwCols; k0 += 8) {
        float v[8];
        tmem_ld8(t_lane + kTmDw + (uint32_t)(t - 1) * kTmDwCols + (uint32_t)k0, v);
        if (live)
          for (int j = 0; j < 8; ++j)
            if (k0 + j < W) dst[(size_t)n * W + k0 + j] = v[j];
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
  if (a.do_grad && tid < a.n_cparams) my_grad[a.n_net_params + tid] = sm.cgS[tid];   // warp 0 -> half-0 row
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------
// dispatch on the jet signature
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

#define TDB_TC_SIGS(X) \
  X(0, 0, 0) X(1, 0, 0) X(2, 0, 0) X(3, 0, 0) X(4, 0, 0) \
  X(1, 1, 0) X(2, 1, 0) X(1, 2, 0) X(2, 2, 0) X(3, 1, 0) X(1, 3, 0) X(3, 2, 0) X(2, 3, 0) X(4, 1, 0) X(1, 4, 0) \
  X(4, 2, 0) X(2, 4, 0) X(3, 3, 0) X(1, 1, 1) X(2, 1, 1) X(2, 2, 1) X(2, 2, 2)

bool jet_tc_supports(int o0, int o1, int o2) {
#define X(A, B, Cc) if (o0 == A && o1 == B && o2 == Cc) return true;
  TDB_TC_SIGS(X)
#undef X
  return false;
}
int jet_tc_points_per_tile(int o0, int o1, int o2) { return kTcParts * (12 / (1 + o0 + o1 + o2)); }
int jet_tc_partial_rows() { return kTcParts; }

cudaError_t launch_jet_tc(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s) {
  const int n_mma = a.n_layers - 2;
#define X(A, B, Cc)                                                            \
  if (o0 == A && o1 == B && o2 == Cc)                                          \
    return n_mma == 1 ? launch_sig<A, B, Cc, 1>(a, wimg, grid, s) : launch_sig<A, B, Cc, 2>(a, wimg, grid, s);
  TDB_TC_SIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace tdb
