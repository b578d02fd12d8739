// Weight-gradient GEMM of the streamed tensor-core path (jet_tcs_kernel.cuh):
//     dW_t[n, k] = sum_r gZ_t[r, n] * Y_{t-1}[r, k]        r = (point, jet channel) columns of the chunk's tiles
// for the W x W layers t = 1..n_mma.  The fused forward / backward kernel streams gZ_t and Y_{t-1} to HBM as fp32
// arrays of float4 [group of 4 columns][Wp neurons]; this kernel is the split-K contraction over the columns:
// HBM-bound (2 * Wp * 4 bytes per column and layer), tcgen05.mma.kind::tf32 in the 3xTF32 split, accumulator in TMEM.
//
// One CTA = one (layer, every splits-th stage); a stage = KB = 32 columns of both operands = two contiguous runs of
// KB * Wp floats.  Three roles:
//   * warp 9 (one lane): cp.async.bulk of the two runs of a stage into a 3-deep ring of raw buffers, completion on an
//     mbarrier - the HBM stream is asynchronous and ~80 KB deep per SM, independent of the registers of the other warps;
//   * warps 0..7: raw buffer -> operand images: hi = tf32 round-to-nearest (integer ALU), lo = x - hi, scalar stores
//     (32 consecutive neurons per warp: conflict-free) into MN-major images (SWIZZLE_128B_BASE32B: [MN block of 32][K
//     rows][32 floats], 32-byte chunks XOR (row & 3)), 2-deep ring;
//   * warp 8 issues, per K-step of 8 rows,
//         D[128 x 112] += gZ_hi^T Y_hi + gZ_hi^T Y_lo + gZ_lo^T Y_hi        (A = gZ image, B = Y image, both MN-major)
//     and releases the image stage with tcgen05.commit.
// The CTA's partial dW goes to its layer's block of partial row blockIdx.x, a row it shares with CTA blockIdx.x of
// jet_tcs_kernel (reduced in a fixed order by reduce_partials_kernel -> bit-reproducible).
#include "jet_tc_kernel.cuh"
#include "jet_tcs.cuh"

namespace tdb {

#ifndef TDB_WG_KB
#define TDB_WG_KB 32
#endif
#ifndef TDB_WG_RAW
#define TDB_WG_RAW 3
#endif
constexpr int kWgKB = TDB_WG_KB;                          // rows (K) per stage: four K-steps (16-row stages with a 10-deep raw
                                                   // ring measured slower: 2.32 vs 1.87 ms on the wave workload)
constexpr int kWgRawStages = TDB_WG_RAW, kWgImgStages = 2;
constexpr int kWgBlock = 4;                        // consecutive stages a CTA takes per round-robin turn
constexpr int kWgImg = 4 * kWgKB * 32;             // floats of one operand image: 4 MN blocks x 16 K rows x 32
constexpr int kWgImgStageFloats = 4 * kWgImg;      // A hi, A lo, B hi, B lo = 64 KB
constexpr int kWgRawOp = kWgKB * 104;              // floats of one operand of a raw stage (Wp <= 104)
constexpr int kWgRawStageFloats = 2 * kWgRawOp;    // 26 KB
constexpr int kWgConvWarps = 8;
constexpr int kWgThreads = (kWgConvWarps + 2) * 32;
constexpr size_t kWgSmemBytes =
    (size_t)(kWgImgStages * kWgImgStageFloats + kWgRawStages * kWgRawStageFloats) * 4 + 1024 + 256;

__device__ __forceinline__ void wg_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_gemm_kernel(const WgradArgs a) {
  extern __shared__ uint8_t smem_raw_wg[];
  const uint32_t s0_ = smem_u32(smem_raw_wg);
  float* const sbase = reinterpret_cast<float*>(smem_raw_wg + (((s0_ + 1023u) & ~1023u) - s0_));
  float* const raw = sbase + kWgImgStages * kWgImgStageFloats;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(raw + kWgRawStages * kWgRawStageFloats);
  uint64_t* const raw_full = bars;                                   // bulk copies -> converters
  uint64_t* const raw_empty = bars + kWgRawStages;                   // converters (8 arrivals) -> producer
  uint64_t* const img_full = bars + 2 * kWgRawStages;                // converters (8 arrivals) -> MMA warp
  uint64_t* const img_empty = img_full + kWgImgStages;               // MMA warp (tcgen05.commit) -> converters
  uint64_t* const done = img_empty + kWgImgStages;
  uint32_t* const tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int layer = blockIdx.x % a.n_mma;          // 0-based: dW of W x W layer `layer + 1`
  const int split = blockIdx.x / a.n_mma;
  // The partial rows are shared with jet_tcs_kernel: its CTA b has zero-filled row b (first chunk of a call) and owns the
  // bias / first-layer / last-layer entries; this CTA b owns the dW block of its layer in the same row - half the rows
  // for reduce_partials_kernel to read and no second zero fill.
  if (split >= a.splits) return;                   // idle CTA
  const int KB = a.kb, Wp = a.Wp;
  const int F = KB * Wp / 4;                       // float4 per operand and stage
  const long long n_kb = (a.total4 + F - 1) / F;
  // Stages are dealt out round-robin in small blocks, not as one contiguous range per CTA: the tensor
  // core ADDS into its fp32 accumulator with truncation, so a long chain of same-signed products (the lambda-weighted
  // boundary rows, which sit together at the end of a stream) drifts by ~2^-24 per MMA - measured 4e-5 on dW when one CTA
  // took all of them.  Interleaved, every CTA sees an equal share of every part of the stream and the partial sums are
  // combined in fp32 round-to-nearest by reduce_partials_kernel.  Each stage is still one contiguous 13 KB run per operand.
  // (blocks of kWgBlock consecutive stages: neighbouring stages keep the DRAM pages of a CTA's reads together)
  const long long n_blk = (n_kb + kWgBlock - 1) / kWgBlock;
  const long long my_blk = split < n_blk ? (n_blk - split + a.splits - 1) / a.splits : 0;
  long long cnt = 0;
  if (my_blk > 0) {
    const long long last = (long long)split + (my_blk - 1) * a.splits;          // this CTA's last block (may be partial)
    const long long in_last = n_kb - last * kWgBlock < kWgBlock ? n_kb - last * kWgBlock : kWgBlock;
    cnt = (my_blk - 1) * kWgBlock + in_last;
  }
  auto stage_of = [&](long long i) { return ((i / kWgBlock) * a.splits + split) * kWgBlock + i % kWgBlock; };
  const float* __restrict__ gs = a.gs + (size_t)layer * a.stream_stride;
  const float* __restrict__ ys = a.ys + (size_t)layer * a.stream_stride;

  // MN pad (neurons >= Wp) of the images stays zero
  for (int i = tid; i < kWgImgStages * kWgImgStageFloats; i += kWgThreads) sbase[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < kWgRawStages; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, kWgConvWarps); }
    for (int s = 0; s < kWgImgStages; ++s) { mbar_init(img_full + s, kWgConvWarps); mbar_init(img_empty + s, 1); }
    mbar_init(done, 1);
  }
  if (warp == kWgConvWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  pdl_wait();                          // the streams (and the zero-filled partial rows) of the fused kernel are complete
  pdl_launch_dependents();

  if (warp == kWgConvWarps + 1) {
    // ---- producer: HBM -> raw ring ------------------------------------------------------------------------------
    if (lane == 0) {
      for (long long i = 0; i < cnt; ++i) {
        const long long kb = stage_of(i);
        const int st = (int)(i % kWgRawStages);
        const uint32_t use = (uint32_t)(i / kWgRawStages);
        if (use > 0) mbar_wait(raw_empty + st, (use - 1) & 1);
        const long long f0 = kb * F;
        const long long nf = a.total4 - f0 < F ? a.total4 - f0 : F;            // float4 of this stage (tail: fewer)
        const uint32_t bytes = (uint32_t)nf * 16u;
        const uint32_t b = smem_u32(raw_full + st);
        float* dst = raw + st * kWgRawStageFloats;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(dst)), "l"(gs + f0 * 4), "r"(bytes), "r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(dst + kWgRawOp)), "l"(ys + f0 * 4), "r"(bytes), "r"(b) : "memory");
      }
    }
  } else if (warp < kWgConvWarps) {
    // ---- converters: thread t owns the float4 t, t + 256, ... of every stage (same image positions every stage) ----
    constexpr int NQ = (kWgKB * 104 / 4 + 255) / 256;      // float4 per thread, operand and stage (4)
    int off[NQ];                                    // image offset of element 0 of the float4 (element e: K row + e)
    int c3[NQ];
    bool own[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int f = tid + 256 * i;
      own[i] = f < F;
      const int n = f % Wp, krow = 4 * (f / Wp);
      off[i] = (n >> 5) * KB * 32 + krow * 32 + (n & 7);
      c3[i] = (n & 31) >> 3;
    }
    auto put = [&](float* hi_img, float* lo_img, int i, const float4& v) {
      const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h = tf32_rna(xv[e]);
        const int o = off[i] + e * 32 + ((c3[i] ^ e) << 3);
        hi_img[o] = h;
        lo_img[o] = xv[e] - h;
      }
    };
    for (long long i = 0; i < cnt; ++i) {
      const long long kb = stage_of(i);
      const int rs = (int)(i % kWgRawStages), is = (int)(i % kWgImgStages);
      const uint32_t ruse = (uint32_t)(i / kWgRawStages), iuse = (uint32_t)(i / kWgImgStages);
      mbar_wait(raw_full + rs, ruse & 1);
      const long long nf = a.total4 - kb * F < F ? a.total4 - kb * F : F;
      const float4* r4 = reinterpret_cast<const float4*>(raw + rs * kWgRawStageFloats);
      float4 gc[NQ], yc[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const bool ok = own[q] && tid + 256 * q < nf;
        gc[q] = ok ? r4[tid + 256 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
        yc[q] = ok ? r4[kWgRawOp / 4 + tid + 256 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (iuse > 0) mbar_wait(img_empty + is, (iuse - 1) & 1);
      float* const st = sbase + is * kWgImgStageFloats;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (own[q]) {
          put(st, st + kWgImg, q, gc[q]);
          put(st + 2 * kWgImg, st + 3 * kWgImg, q, yc[q]);
        }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { wg_arrive(raw_empty + rs); wg_arrive(img_full + is); }   // raw stage consumed: refill it
    }
  } else {
    // ---- MMA warp ---------------------------------------------------------------------------------------------
    constexpr uint32_t idesc = umma_idesc(128, 112, 1, 1);
    int stage = 0;
    uint32_t fph = 0;
    const bool leader = elect_one();
    for (long long i = 0; i < cnt; ++i) {
      mbar_wait(img_full + stage, fph);
      tc_fence_after();
      const float* st = sbase + stage * kWgImgStageFloats;
      // MN blocks of 32 at LBO = one block (KB rows x 128 B); 8 K rows = two 4-row swizzle atoms (SBO = 512 B)
      const uint64_t ah = umma_desc(smem_u32(st), KB * 128, 512, 1), al = umma_desc(smem_u32(st + kWgImg), KB * 128, 512, 1);
      const uint64_t bh = umma_desc(smem_u32(st + 2 * kWgImg), KB * 128, 512, 1),
                     bl = umma_desc(smem_u32(st + 3 * kWgImg), KB * 128, 512, 1);
      for (int s = 0; s < KB / 8; ++s) {
        const uint64_t o = ((uint64_t)s * 1024) >> 4;
        if (leader) {
          umma_tf32(tmem, ah + o, bh + o, idesc, (i > 0 || s > 0) ? 1u : 0u);
          umma_tf32(tmem, ah + o, bl + o, idesc, 1u);
          umma_tf32(tmem, al + o, bh + o, idesc, 1u);
        }
      }
      if (leader) umma_commit(img_empty + stage);
      __syncwarp();
      if (++stage == kWgImgStages) { stage = 0; fph ^= 1; }
    }
    if (leader) umma_commit(done);
    __syncwarp();
  }

  // ---- epilogue: TMEM -> shared memory (transposition buffer) -> this CTA's partial row, coalesced -----------------
  float* const tbuf = sbase;                       // [128][113] floats: the image stages are free by now
  if (warp < kWgConvWarps) {
    const int n = (warp & 3) * 32 + lane, half = warp >> 2;
    if (cnt > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    for (int k0 = half * 64; k0 < half * 64 + 64 && k0 < 112; k0 += 16) {
      float v[16];
      if (cnt > 0) tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)k0, v);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) tbuf[n * 113 + k0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  {
    const int W = a.W;
    float* const dst = a.part + (size_t)blockIdx.x * a.n_params_pad + a.w_off[layer];
    for (int i = tid; i < W * W; i += kWgThreads) {
      const int n = i / W, k = i - n * W;
      const float v = tbuf[n * 113 + k];
      dst[i] = a.accumulate ? dst[i] + v : v;
    }
  }
  if (warp == kWgConvWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem) : "memory");
}

int wgrad_kb() { return kWgKB; }

cudaError_t launch_wgrad_gemm(const WgradArgs& a, int grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  return launch_pdl(wgrad_gemm_kernel, dim3(grid), dim3(kWgThreads), kWgSmemBytes, s, a);
}

}  // namespace tdb
