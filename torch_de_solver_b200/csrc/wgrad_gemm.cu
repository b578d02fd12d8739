// Weight-gradient GEMM of the streamed tensor-core path (jet_tcs_kernel.cuh):
//     dW_t[n, k] = sum_r gZ_t[r, n] * Y_{t-1}[r, k]        r = (point, jet channel) rows of the chunk
// for the W x W layers t = 1..n_mma.  The fused forward / backward kernel streams gZ_t and Y_{t-1} as fp32 row
// arrays [rows][Wp] to HBM; this kernel is the split-K contraction over the rows: HBM-bound (2 * Wp * 4 bytes per
// row and layer), tcgen05.mma.kind::tf32 in the 3xTF32 split with the accumulator in TMEM.
//
// One CTA = one (layer, row range).  Warps 0..7 stream the rows: coalesced 16-byte loads (one row = Wp / 4 lanes),
// hi = cvt.rna.tf32, lo = x - hi, 16-byte stores into MN-major operand images (SWIZZLE_128B_BASE32B: [MN block of
// 32][K rows][32 floats], 32-byte chunks XOR (row & 3)) of a 3-stage ring; warp 8 issues, per K-step of 8 rows,
//     D[128 x 112] += gZ_hi^T Y_hi + gZ_hi^T Y_lo + gZ_lo^T Y_hi        (A = gZ image, B = Y image, both MN-major)
// and releases the stage with tcgen05.commit.  The CTA's partial dW goes to its own row of the gradient partial
// buffer (reduced in a fixed order by reduce_partials_kernel -> bit-reproducible).
#include "jet_tc_kernel.cuh"
#include "jet_tcs.cuh"

namespace tdb {

constexpr int kWgKB = 32;                          // rows (K) per stage
constexpr int kWgStages = 3;
constexpr int kWgImg = 4 * kWgKB * 32;             // floats of one operand image: 4 MN blocks x 32 K rows x 32
constexpr int kWgStageFloats = 4 * kWgImg;         // A hi, A lo, B hi, B lo = 64 KB
constexpr int kWgLoaderWarps = 8;
constexpr int kWgThreads = (kWgLoaderWarps + 1) * 32;
constexpr size_t kWgSmemBytes = (size_t)kWgStages * kWgStageFloats * 4 + 1024 + 128;

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_gemm_kernel(const WgradArgs a) {
  extern __shared__ uint8_t smem_raw_wg[];
  const uint32_t s0_ = smem_u32(smem_raw_wg);
  float* const sbase = reinterpret_cast<float*>(smem_raw_wg + (((s0_ + 1023u) & ~1023u) - s0_));
  uint64_t* const bars = reinterpret_cast<uint64_t*>(sbase + kWgStages * kWgStageFloats);
  uint64_t* const full = bars;                     // [stages] loader warps -> MMA warp
  uint64_t* const empty = bars + kWgStages;        // [stages] MMA warp (tcgen05.commit) -> loader warps
  uint64_t* const done = bars + 2 * kWgStages;
  uint32_t* const tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int layer = blockIdx.x % a.n_mma;          // 0-based: dW of W x W layer `layer + 1`
  const int split = blockIdx.x / a.n_mma;
  if (!a.accumulate)                               // first chunk of a call: the row holds nothing but this CTA's dW block
    for (int i = tid; i < a.n_params_pad; i += kWgThreads) a.part[(size_t)blockIdx.x * a.n_params_pad + i] = 0.f;
  if (split >= a.splits) return;                   // idle CTA: its partial row stays zero
  const long long n_kb = (a.rows + kWgKB - 1) / kWgKB;
  const long long kb0 = n_kb * split / a.splits, kb1 = n_kb * (split + 1) / a.splits;
  const float* __restrict__ gs = a.gs + (size_t)layer * a.stream_stride;
  const float* __restrict__ ys = a.ys + (size_t)layer * a.stream_stride;
  const int Wp = a.Wp, nq = Wp >> 2;               // float4 per row

  for (int i = tid; i < kWgStages * kWgStageFloats; i += kWgThreads) sbase[i] = 0.f;     // MN pad (>= Wp) stays zero
  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full + s, kWgLoaderWarps); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
  }
  if (warp == kWgLoaderWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp < kWgLoaderWarps) {
    // ---- loaders: warp w owns K rows 4w .. 4w + 3 of every stage --------------------------------------------
    float4 g[4], y[4];
    auto fetch = [&](long long kb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long r = kb * kWgKB + warp * 4 + i;
        const bool ok = lane < nq && r < a.rows;
        g[i] = ok ? __ldcs(reinterpret_cast<const float4*>(gs + (size_t)r * Wp) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        y[i] = ok ? __ldcs(reinterpret_cast<const float4*>(ys + (size_t)r * Wp) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto put = [&](float* hi_img, float* lo_img, int krow, const float4& v) {
      const float x[4] = {v.x, v.y, v.z, v.w};
      float h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[j]));
        h[j] = __uint_as_float(hb);
        l[j] = x[j] - h[j];
      }
      const int off = sw_off_mn(krow, 4 * lane, kWgKB);
      st4(hi_img + off, h);
      st4(lo_img + off, l);
    };
    if (kb0 < kb1) fetch(kb0);
    for (long long kb = kb0; kb < kb1; ++kb) {
      const long long i = kb - kb0;
      const int stage = (int)(i % kWgStages);
      const uint32_t use = (uint32_t)(i / kWgStages);           // how often this stage was filled before
      if (use > 0) mbar_wait(empty + stage, (use - 1) & 1);     // ... and its (use)-th release by the MMA warp
      float* const st = sbase + stage * kWgStageFloats;
      float4 gc[4], yc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { gc[q] = g[q]; yc[q] = y[q]; }
      if (kb + 1 < kb1) fetch(kb + 1);              // next block's loads fly while this one is converted
      if (lane < nq) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          put(st, st + kWgImg, warp * 4 + q, gc[q]);
          put(st + 2 * kWgImg, st + 3 * kWgImg, warp * 4 + q, yc[q]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(full + stage)) : "memory");
    }
  } else {
    // ---- MMA warp ---------------------------------------------------------------------------------------------
    constexpr uint32_t idesc = umma_idesc(128, 112, 1, 1);
    int stage = 0;
    uint32_t fph = 0;
    const bool leader = elect_one();
    for (long long kb = kb0; kb < kb1; ++kb) {
      mbar_wait(full + stage, fph);
      tc_fence_after();
      const float* st = sbase + stage * kWgStageFloats;
      // MN blocks of 32 at LBO = one block (kWgKB rows x 128 B); 8 K rows = two 4-row swizzle atoms (SBO = 512 B)
      const uint64_t ah = umma_desc(smem_u32(st), kWgKB * 128, 512, 1), al = umma_desc(smem_u32(st + kWgImg), kWgKB * 128, 512, 1);
      const uint64_t bh = umma_desc(smem_u32(st + 2 * kWgImg), kWgKB * 128, 512, 1),
                     bl = umma_desc(smem_u32(st + 3 * kWgImg), kWgKB * 128, 512, 1);
#pragma unroll
      for (int s = 0; s < kWgKB / 8; ++s) {
        const uint64_t o = ((uint64_t)s * 1024) >> 4;
        if (leader) {
          umma_tf32(tmem, ah + o, bh + o, idesc, (kb > kb0 || s > 0) ? 1u : 0u);
          umma_tf32(tmem, ah + o, bl + o, idesc, 1u);
          umma_tf32(tmem, al + o, bh + o, idesc, 1u);
        }
      }
      if (leader) umma_commit(empty + stage);
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; fph ^= 1; }
    }
    if (leader) umma_commit(done);
    __syncwarp();
  }

  // ---- epilogue: TMEM -> this CTA's partial row --------------------------------------------------------------
  if (warp < kWgLoaderWarps) {
    const int W = a.W, n = (warp & 3) * 32 + lane, half = warp >> 2;
    float* const dst = a.part + (size_t)blockIdx.x * a.n_params_pad + a.w_off[layer];
    if (kb0 < kb1) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    for (int k0 = half * 64; k0 < half * 64 + 64 && k0 < 112; k0 += 16) {
      float v[16];
      if (kb0 < kb1) tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)k0, v);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (n < W) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (k0 + j < W) {
            float* q = dst + (size_t)n * W + k0 + j;
            *q = a.accumulate ? *q + v[j] : v[j];
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWgLoaderWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem) : "memory");
}

cudaError_t launch_wgrad_gemm(const WgradArgs& a, int grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  wgrad_gemm_kernel<<<grid, kWgThreads, kWgSmemBytes, s>>>(a);
  return cudaGetLastError();
}

}  // namespace tdb
