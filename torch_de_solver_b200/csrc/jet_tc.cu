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

constexpr int kTcThreads = 256;
constexpr int kTcCols = 48;                      // (point, channel) columns per tile = MMA N
constexpr int kTcWRows = 104;                    // rows of the weight image (neurons padded to 8)
constexpr int kTcActBlock = kTcCols * 32;        // floats per k-block of an activation operand
constexpr int kTcActFloats = 4 * kTcActBlock;    // 6144 floats = 24 KB
constexpr int kTcWBlock = kTcWRows * 32;
constexpr int kTcWFloats = 4 * kTcWBlock;        // 13312 floats = 52 KB
constexpr int kTcMaxMma = 3;                     // W x W layers whose dW fits TMEM (64 + 3 * 128 <= 512 columns)
constexpr int kTcSavePitch = 104;
// TMEM columns: D accumulator | gZ hi | gZ lo (A operands of the weight-gradient MMA) | dW slots (112 each)
constexpr uint32_t kTmD = 0, kTmAHi = 64, kTmALo = 120, kTmDw = 176, kTmDwCols = 112;

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
// D[128 x 48] (+)= A(W image, K-major) . B(act, K-major), 3xTF32
__device__ __forceinline__ void issue_forward(uint32_t d_tmem, const float* w_hi, const float* w_lo,
                                              const float* b_hi, const float* b_lo, int ksteps) {
  constexpr uint32_t idesc = umma_idesc(128, kTcCols, 0, 0);
  const uint32_t wa = smem_u32(w_hi), wl = smem_u32(w_lo), ba = smem_u32(b_hi), bl = smem_u32(b_lo);
  uint32_t acc = 0;
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t A = pass == 0 ? wl : wa;          // lo*hi, hi*lo, hi*hi
    const uint32_t B = pass == 1 ? bl : ba;
    for (int s = 0; s < ksteps; ++s) {
      const uint32_t ao = (uint32_t)(s >> 2) * kTcWBlock * 4 + (uint32_t)(s & 3) * 32;
      const uint32_t bo = (uint32_t)(s >> 2) * kTcActBlock * 4 + (uint32_t)(s & 3) * 32;
      umma_tf32(d_tmem, umma_desc(A + ao, 16, 1024), umma_desc(B + bo, 16, 1024), idesc, acc);
      acc = 1;
    }
  }
}
// dW[128 (n) x 112 (k)] += A(gZ in TMEM: lanes n, columns (pc)) . B(Y in smem read MN-major: N = k, K = (pc))
__device__ __forceinline__ void issue_wgrad(uint32_t d_tmem, uint32_t a_hi_tmem, uint32_t a_lo_tmem,
                                            const float* y_hi, const float* y_lo, uint32_t accumulate) {
  constexpr uint32_t idesc = umma_idesc(128, 112, 0, 1);
  const uint32_t ya = smem_u32(y_hi), yl = smem_u32(y_lo);
  uint32_t acc = accumulate;
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t A = pass == 0 ? a_lo_tmem : a_hi_tmem;
    const uint32_t B = pass == 1 ? yl : ya;
    for (int s = 0; s < kTcCols / 8; ++s) {
      // 8 (pc) rows = two 4-row swizzle atoms (SBO = 512 B); MN blocks of 32 k at LBO = one k-block
      umma_tf32_ts(d_tmem, A + (uint32_t)s * 8, umma_desc(B + (uint32_t)s * 1024, kTcActBlock * 4, 512, 1), idesc, acc);
      acc = 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct TcSmem {
  float *w_hi, *w_lo, *a_hi, *a_lo, *b_hi, *b_lo;
  float *xS, *uS, *guS, *wlS, *cgS;
  double* lossS;
  uint64_t* bar;
  uint32_t* tmem_ptr;
};
constexpr size_t kTcSmemBytes = (size_t)(2 * kTcWFloats + 4 * kTcActFloats) * 4 + 1024 /*align*/ +
                                (kTcCols * 4 + 2 * kMaxOut * kTcCols + kMaxOut * kTcSavePitch + kMaxCParams) * 4 +
                                32 * 8 + 64;

size_t jet_tc_smem_bytes() { return kTcSmemBytes; }

__device__ __forceinline__ void load_w_image(const TcSmem& sm, const float* __restrict__ img) {
  const float4* s4 = reinterpret_cast<const float4*>(img);
  float4* d4 = reinterpret_cast<float4*>(sm.w_hi);               // w_hi and w_lo are contiguous
  for (int i = threadIdx.x; i < 2 * kTcWFloats / 4; i += kTcThreads) d4[i] = __ldg(s4 + i);
}

__global__ void __launch_bounds__(kTcThreads, 1) jet_tc_kernel(const JetArgs a, const float* __restrict__ wimg) {
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
    sm.xS = f; f += kTcCols * 4;
    sm.uS = f; f += kMaxOut * kTcCols;
    sm.guS = f; f += kMaxOut * kTcCols;
    sm.wlS = f; f += kMaxOut * kTcSavePitch;
    sm.cgS = f; f += kMaxCParams;
    sm.lossS = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(f) + 15) & ~uintptr_t(15));
    sm.bar = reinterpret_cast<uint64_t*>(sm.lossS + 32);
    sm.tmem_ptr = reinterpret_cast<uint32_t*>(sm.bar + 1);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = (warp & 3) * 32 + lane;                 // neuron = TMEM lane owned by this thread
  const int half = warp >> 2;                           // which half of the tile's points
  const int L = a.n_layers, W = a.widths[1], n_out = a.widths[L], d = a.d;
  const int n_mma = L - 2;
  const int ksteps = (W + 7) / 8;
  const bool live = n < W;
  // two gradient-partial rows per CTA (one per point half): every address has a single owner thread, so the
  // fp32 accumulation order is fixed and results are bit-reproducible
  float* const my_grad = a.part_grad + ((size_t)blockIdx.x * 2 + half) * a.n_params_pad;
  float* const my_scratch = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
  const size_t save_block = (size_t)kTcCols * kTcSavePitch;          // one saved [48][104] block

  // ---- one-time setup --------------------------------------------------------------------------------
  for (int i = tid; i < 2 * a.n_params_pad; i += kTcThreads)
    a.part_grad[(size_t)blockIdx.x * 2 * a.n_params_pad + i] = 0.f;
  for (int i = tid; i < 4 * kTcActFloats; i += kTcThreads) sm.a_hi[i] = 0.f;      // pad rows / columns stay zero
  if (tid < 32) sm.lossS[tid] = 0.0;
  if (tid < kMaxCParams) sm.cgS[tid] = 0.f;
  for (int i = tid; i < n_out * W; i += kTcThreads) sm.wlS[(i / W) * kTcSavePitch + i % W] = a.arena[a.w_off[L - 1] + i];
  if (tid == 0) mbar_init(sm.bar, 1);
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
      for (uint32_t c = kTmAHi; c < kTmDw; c += 8) tmem_st_n(t_lane + c, z8, 8);
    tmem_st_wait();
  }
  uint32_t phase = 0;
  uint32_t dw_started = 0;
  int resident = 0;                                     // W x W layer whose image is in shared memory

  int seg_i = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    while (tile >= a.seg_tile_begin[seg_i + 1]) ++seg_i;
    const tdb200_segment& sg = a.segs[seg_i];
    const int ndirs = sg.n_dirs, ncols = sg.n_cols;
    int J = 1;
    for (int i = 0; i < ndirs; ++i) J += sg.dir_order[i];
    const int P = kTcCols / J;
    const int Ph = (P + 1) / 2;
    const int p_lo = half == 0 ? 0 : Ph, p_hi = half == 0 ? Ph : P;
    const long long g_first = (long long)(tile - a.seg_tile_begin[seg_i]) * P;
    const int p_valid = (int)min((long long)P, sg.n_groups - g_first);

    for (int i = tid; i < P * d; i += kTcThreads) {
      const int p = i / d, ax = i - p * d;
      sm.xS[p * 4 + ax] = p < p_valid ? __ldg(a.pts + (size_t)(sg.pts_off + g_first + p) * d + ax) : 0.f;
    }
    if (resident != 2) { load_w_image(sm, wimg); resident = 2; }   // W of layer 1 (key = 2 * layer + transposed)
    __syncthreads();

    // ---- layer 0 (K = d): thread-local ----------------------------------------------------------------
    if (live) {
      const float* W0 = a.arena + a.w_off[0];
      const float b0 = a.arena[a.b_off[0] + n];
      float w0[4];
      for (int ax = 0; ax < d; ++ax) w0[ax] = W0[n * d + ax];
      float* ysave = my_scratch;
      for (int p = p_lo; p < p_hi; ++p) {
        float z0 = b0;
        for (int ax = 0; ax < d; ++ax) z0 = fmaf(w0[ax], sm.xS[p * 4 + ax], z0);
        const float av = tanhf(z0);
        const TanhF f(av);
        int r = p * J;
        split_store(sm.a_hi, sm.a_lo, sw_off(r, n, kTcCols), av);
        ysave[(size_t)r * kTcSavePitch + n] = av;
        int c = 1;
        for (int i = 0; i < ndirs; ++i) {
          const int o = sg.dir_order[i];
          float z[4] = {w0[sg.dir_axis[i]], 0.f, 0.f, 0.f}, y[4];
          tanh_jet_fwd(f, z, o, y);
          for (int k = 0; k < o; ++k) {
            split_store(sm.a_hi, sm.a_lo, sw_off(r + c + k, n, kTcCols), y[k]);
            ysave[(size_t)(r + c + k) * kTcSavePitch + n] = y[k];
          }
          c += o;
        }
      }
    }

    // ---- W x W layers: tensor-core GEMM + thread-local tanh-jet epilogue ------------------------------
    for (int l = 1; l <= n_mma; ++l) {
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        issue_forward(tmem + kTmD, sm.w_hi, sm.w_lo, sm.a_hi, sm.a_lo, ksteps);
        umma_commit(sm.bar);
      }
      mbar_wait(sm.bar, phase);
      phase ^= 1;
      tc_fence_after();
      // the weights of the next GEMM (layer l + 1 forward, or layer n_mma again for the backward sweep) can be
      // fetched while the epilogue runs: the tensor core is done with the image
      if (l < n_mma) { load_w_image(sm, wimg + (size_t)l * 4 * kTcWFloats); resident = 2 * (l + 1); }
      float* ysave = my_scratch + (size_t)(2 * l) * save_block;
      float* zsave = ysave + save_block;
      const float bl = live ? a.arena[a.b_off[l] + n] : 0.f;
      for (int p = p_lo; p < p_hi; ++p) {
        float zc[8];
        tmem_ld8(t_lane + kTmD + (uint32_t)(p * J), zc);
        if (!live) continue;
        const float av = tanhf(zc[0] + bl);
        const TanhF f(av);
        const int r = p * J;
        split_store(sm.a_hi, sm.a_lo, sw_off(r, n, kTcCols), av);
        ysave[(size_t)r * kTcSavePitch + n] = av;
        zsave[(size_t)r * kTcSavePitch + n] = av;
        int c = 1;
        for (int i = 0; i < ndirs; ++i) {
          const int o = sg.dir_order[i];
          float y[4];
          tanh_jet_fwd(f, zc + c, o, y);
          for (int k = 0; k < o; ++k) {
            split_store(sm.a_hi, sm.a_lo, sw_off(r + c + k, n, kTcCols), y[k]);
            ysave[(size_t)(r + c + k) * kTcSavePitch + n] = y[k];
            zsave[(size_t)(r + c + k) * kTcSavePitch + n] = zc[c + k];
          }
          c += o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();

    // ---- last layer (n_out <= 8 outputs): u[v][r] = sum_n Wl[v][n] Y[r][n] -----------------------------
    const int PJ = P * J;
    for (int idx = tid; idx < n_out * PJ; idx += kTcThreads) {
      const int v = idx / PJ, r = idx - v * PJ;
      float s = (r % J) == 0 ? a.arena[a.b_off[L - 1] + v] : 0.f;
      const float* wl = sm.wlS + v * kTcSavePitch;
      for (int k = 0; k < W; ++k) {
        const int o = sw_off(r, k, kTcCols);
        s = fmaf(wl[k], sm.a_hi[o] + sm.a_lo[o], s);
      }
      sm.uS[v * kTcCols + r] = s;
      sm.guS[v * kTcCols + r] = 0.f;
    }
    __syncthreads();

    // ---- operator terms, residual, loss, adjoint seeds (one thread per point) -------------------------
    if (tid < p_valid) {
      const int p = tid;
      const long long row = g_first + p;
      for (int col = 0; col < ncols; ++col) {
        float val = 0.f;
        for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
          const tdb200_term tm = a.terms[t];
          float prod = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                               : a.arena[a.n_net_params + tm.idx];
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = a.factors[fi];
            prod *= pow_i(sm.uS[fc.var * kTcCols + p * J + fc.chan], fc.ipow, fc.pow);
          }
          val += prod;
        }
        if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
        const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
        const float res = val - tgt;
        const int slot = sg.col_slot[col];
        atomicAdd(&sm.lossS[slot], (double)res * (double)res);
        if (!a.do_grad) continue;
        const float seed = 2.f * __ldg(a.slot_scale + slot) * res;
        for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
          const tdb200_term tm = a.terms[t];
          const float cf = tm.kind == 0 ? tm.coeff : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                                                   : a.arena[a.n_net_params + tm.idx];
          float full = 1.f;
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = a.factors[fi];
            const float x = sm.uS[fc.var * kTcCols + p * J + fc.chan];
            float part = seed * cf * dpow_i(x, fc.ipow, fc.pow);
            for (int fj = tm.fac_begin; fj < tm.fac_end; ++fj) {
              if (fj == fi) continue;
              const tdb200_factor fo = a.factors[fj];
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
    if (!a.do_grad) continue;

    // ---- backward of the last layer: dWl, dbl (thread-local partial sums over this thread's points) ----
    if (tid < n_out) {                                   // warp 0 -> half 0 row
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += sm.guS[tid * kTcCols + p * J];
      atomicAdd(my_grad + a.b_off[L - 1] + tid, s);
    }
    if (live) {
      const float* ylast = my_scratch + (size_t)(2 * n_mma) * save_block;      // Y of the last hidden layer
      for (int v = 0; v < n_out; ++v) {
        float s = 0.f;
        for (int r = p_lo * J; r < p_hi * J; ++r) s = fmaf(sm.guS[v * kTcCols + r], ylast[(size_t)r * kTcSavePitch + n], s);
        atomicAdd(my_grad + a.w_off[L - 1] + v * W + n, s);
      }
    }

    // ---- backward sweep over the tanh layers t = n_mma .. 0 ------------------------------------------
    float dw0[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = n_mma; t >= 0; --t) {
      const float* ysave = my_scratch + (size_t)(2 * t) * save_block;
      const float* zsave = ysave + save_block;
      float db = 0.f;
      float w0[4] = {0.f, 0.f, 0.f, 0.f};
      if (t == 0 && live)
        for (int ax = 0; ax < d; ++ax) w0[ax] = a.arena[a.w_off[0] + n * d + ax];
      if (t > 0 && resident != 2 * t + 1) {            // W_t^T for the backward-data GEMM of this layer
        load_w_image(sm, wimg + (size_t)(t - 1) * 4 * kTcWFloats + 2 * kTcWFloats);
        resident = 2 * t + 1;
      }
      for (int p = p_lo; p < p_hi; ++p) {
        const int r = p * J;
        float gy[8], gzv[8], ghi[8], glo[8];
        if (t == n_mma) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float s = 0.f;
            if (c < J && live)
              for (int v = 0; v < n_out; ++v) s = fmaf(sm.wlS[v * kTcSavePitch + n], sm.guS[v * kTcCols + r + c], s);
            gy[c] = s;
          }
        } else {
          tmem_ld8(t_lane + kTmD + (uint32_t)r, gy);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) gzv[c] = 0.f;
        if (live) {
          const float av = ysave[(size_t)r * kTcSavePitch + n];
          const TanhF f(av);
          float g0 = gy[0] * f.f1;
          int c = 1;
          for (int i = 0; i < ndirs; ++i) {
            const int o = sg.dir_order[i];
            float z[4] = {0.f, 0.f, 0.f, 0.f}, gz[4];
            if (t == 0) z[0] = w0[sg.dir_axis[i]];
            else for (int k = 0; k < o; ++k) z[k] = zsave[(size_t)(r + c + k) * kTcSavePitch + n];
            g0 += tanh_jet_bwd(f, z, gy + c, o, gz);
            if (t == 0) dw0[sg.dir_axis[i]] += gz[0];
            for (int k = 0; k < o; ++k) gzv[c + k] = gz[k];
            c += o;
          }
          gzv[0] = g0;
          db += g0;
          if (t == 0) for (int ax = 0; ax < d; ++ax) dw0[ax] = fmaf(g0, sm.xS[p * 4 + ax], dw0[ax]);
        }
        if (t > 0) {
          // gZ goes to shared memory (B operand of the backward-data GEMM, K-major over n) and to TMEM
          // (A operand of the weight-gradient GEMM: lanes n, columns (pc)), both as hi / lo tf32 pairs
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(gzv[c]));
            ghi[c] = __uint_as_float(hb);
            glo[c] = gzv[c] - ghi[c];
          }
          if (live)
            for (int c = 0; c < J; ++c) {
              const int o = sw_off(r + c, n, kTcCols);
              sm.b_hi[o] = ghi[c];
              sm.b_lo[o] = glo[c];
            }
          tmem_st_n(t_lane + kTmAHi + (uint32_t)r, ghi, J);
          tmem_st_n(t_lane + kTmALo + (uint32_t)r, glo, J);
        }
      }
      if (live) atomicAdd(my_grad + a.b_off[t] + n, db);
      if (t == 0) {
        if (live) for (int ax = 0; ax < d; ++ax) atomicAdd(my_grad + a.w_off[0] + n * d + ax, dw0[ax]);
        break;
      }
      tmem_st_wait();
      // Y_{t-1} (all channels) back from scratch as the MN-major B operand of the weight-gradient GEMM
      {
        const float* yprev = my_scratch + (size_t)(2 * (t - 1)) * save_block;
        for (int idx = tid; idx < PJ * kTcSavePitch; idx += kTcThreads) {
          const int r = idx / kTcSavePitch, k = idx - r * kTcSavePitch;
          if (k < W) split_store(sm.a_hi, sm.a_lo, sw_off_mn(r, k, kTcCols), yprev[idx]);
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        issue_wgrad(tmem + kTmDw + (uint32_t)(t - 1) * kTmDwCols, tmem + kTmAHi, tmem + kTmALo, sm.a_hi, sm.a_lo,
                    dw_started);
        issue_forward(tmem + kTmD, sm.w_hi, sm.w_lo, sm.b_hi, sm.b_lo, ksteps);   // A = W_t^T image
        umma_commit(sm.bar);
      }
      mbar_wait(sm.bar, phase);
      phase ^= 1;
      tc_fence_after();
    }
    dw_started = 1;
    tc_fence_before();
    __syncthreads();
  }

  // ---- flush: dW accumulators (TMEM) and per-CTA scalars ----------------------------------------------
  __syncthreads();
  tc_fence_after();
  if (a.do_grad && dw_started) {
    float* row0 = a.part_grad + (size_t)blockIdx.x * 2 * a.n_params_pad;       // dW lives in the half-0 row
    for (int t = 1; t <= n_mma; ++t) {
      float* dst = row0 + a.w_off[t];
      for (int k0 = half * 56; k0 < half * 56 + 56; k0 += 8) {
        float v[8];
        tmem_ld8(t_lane + kTmDw + (uint32_t)(t - 1) * kTmDwCols + (uint32_t)k0, v);
        if (live)
          for (int j = 0; j < 8; ++j)
            if (k0 + j < W) dst[(size_t)n * W + k0 + j] = v[j];
      }
    }
  }
  if (tid < a.n_slots) a.part_loss[(size_t)blockIdx.x * a.n_slots + tid] = sm.lossS[tid];
  if (a.do_grad && tid < a.n_cparams) my_grad[a.n_net_params + tid] = sm.cgS[tid];   // warp 0 -> half 0 row
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

cudaError_t launch_jet_tc(const JetArgs& a, const float* wimg, int grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(jet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  jet_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, s>>>(a, wimg);
  return cudaGetLastError();
}

}  // namespace tdb
