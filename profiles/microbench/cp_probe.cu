// Layout / rate probe for tcgen05.cp (shared memory -> tensor memory) on sm_100a (not product code; evidence for DESIGN.md).
//   1. layout: a K-major SWIZZLE_128B image of a [128 x 128] fp32 matrix whose element (n, k) holds the number n * 128 + k is
//      copied K-step by K-step (128 rows x 256 bits) into TMEM with the SAME descriptor a kind::tf32 SS MMA uses for that
//      K-step; the TMEM columns are read back and decoded: which (n, k) sits in (lane, column)?
//   2. rate: cycles per "layer GEMM" (13 K-steps x 3 MMAs, N = 64) for SS operands against [26 tcgen05.cp + TS operands],
//      with two accumulators (two tile slots) sharing one copy of the weights.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cp_probe cp_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)type << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
static inline int sw_off(int row, int k, int rows) {       // K-major SWIZZLE_128B
  return (k >> 5) * rows * 32 + row * 32 + ((((k & 31) >> 2) ^ (row & 7)) << 2) + (k & 3);
}

// ---- 1. layout: copy 13 K-steps, read 104 TMEM columns back --------------------------------------------------------
__global__ void __launch_bounds__(128) probe_cp_layout(const float* __restrict__ Aimg, float* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  float* A = reinterpret_cast<float*>(base);                 // [4][128][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 65536);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4 * 128 * 32; i += 128) A[i] = Aimg[i];
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    const bool leader = elect_one();
    for (int s = 0; s < 13; ++s) {
      const uint32_t ao = (uint32_t)(s >> 2) * 128 * 128 + (uint32_t)(s & 3) * 32;
      const uint64_t da = umma_desc(smem_u32(A) + ao, 16, 1024, 2);
      if (leader) tmem_cp_128x256b(tmem + (uint32_t)s * 8, da);
    }
    if (leader) umma_commit(bar);
    __syncwarp();
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 104; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[tid * 104 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem) : "memory");
}

// ---- 2. rate: one warp; per repetition one "layer": weights for two tile slots ---------------------------------------
// MODE 0: SS, 2 slots x 39 MMAs.  MODE 1: 26 cp (hi + lo image) interleaved per K-step with the slot-0 MMAs (TS), then
// the 39 TS MMAs of slot 1.
template <int MODE>
__global__ void __launch_bounds__(32) probe_cp_rate(int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  float* Ah = reinterpret_cast<float*>(base);                 // 52 KB
  float* Al = reinterpret_cast<float*>(base + 53248);
  float* B = reinterpret_cast<float*>(base + 106496);         // [hi | lo] x 2 slots: 4 x 26 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 106496 + 4 * 26624);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (106496 + 4 * 26624) / 4; i += 32) Ah[i] = 0.f;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tptr)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  constexpr uint32_t idesc = idesc_tf32(128, 64, 0, 1);
  const uint64_t dah = umma_desc(smem_u32(Ah), 16, 1024, 2), dal = umma_desc(smem_u32(Al), 16, 1024, 2);
  const bool leader = elect_one();
  const uint32_t t_wh = tmem + 128, t_wl = tmem + 128 + 104;
  uint32_t phase = 0;
  long long t0 = 0;
  for (int r = -2; r < reps; ++r) {
    if (r == 0) t0 = clock64();
    for (int slot = 0; slot < 2; ++slot) {
      const uint64_t dbh = umma_desc(smem_u32(B) + slot * 2 * 26624, 104 * 128, 512, 1),
                     dbl = umma_desc(smem_u32(B) + slot * 2 * 26624 + 26624, 104 * 128, 512, 1);
      const uint32_t d = tmem + slot * 64;
#pragma unroll
      for (int s = 0; s < 13; ++s) {
        const uint64_t ao = ((uint64_t)(s >> 2) * 104 * 128 + (uint64_t)(s & 3) * 32) >> 4;
        const uint64_t bo = ((uint64_t)s * 1024) >> 4;
        if (leader) {
          if (MODE == 0) {
            mma_tf32_ss(d, dah + ao, dbh + bo, idesc, s ? 1u : 0u);
            mma_tf32_ss(d, dah + ao, dbl + bo, idesc, 1u);
            mma_tf32_ss(d, dal + ao, dbh + bo, idesc, 1u);
          } else {
            if (slot == 0) { tmem_cp_128x256b(t_wh + s * 8, dah + ao); tmem_cp_128x256b(t_wl + s * 8, dal + ao); }
            mma_tf32_ts(d, t_wh + s * 8, dbh + bo, idesc, s ? 1u : 0u);
            mma_tf32_ts(d, t_wh + s * 8, dbl + bo, idesc, 1u);
            mma_tf32_ts(d, t_wl + s * 8, dbh + bo, idesc, 1u);
          }
        }
      }
    }
    if (leader) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0) {
    std::vector<float> img(4 * 128 * 32, -1.f);
    for (int n = 0; n < 128; ++n)
      for (int k = 0; k < 128; ++k) img[sw_off(n, k, 128)] = (float)(n * 128 + k);
    float *dA, *dO;
    CK(cudaMalloc(&dA, img.size() * 4));
    CK(cudaMalloc(&dO, 128 * 104 * 4));
    CK(cudaMemcpy(dA, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_cp_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024 + 64));
    probe_cp_layout<<<1, 128, 65536 + 1024 + 64>>>(dA, dO);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(128 * 104);
    CK(cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int lane = 0; lane < 128; ++lane)
      for (int c = 0; c < 104; ++c) {
        const int v = (int)o[lane * 104 + c];
        if (v != lane * 128 + c) {
          if (bad < 24) printf("  TMEM(lane %3d, col %3d) holds A(n = %d, k = %d), expected (%d, %d)\n", lane, c, v / 128, v % 128, lane, c);
          ++bad;
        }
      }
    printf("cp layout: %d mismatches of %d (0 = the SS descriptor of a K-step places A(n, k) at lane n, column k)\n", bad, 128 * 104);
    return 0;
  }
  long long* d_out;
  CK(cudaMalloc(&d_out, 8 * 256));
  const size_t smem = 106496 + 4 * 26624 + 1024 + 64;
  const int reps = 50;
  for (int grid : {1, 148}) {
    long long h[148];
    CK(cudaFuncSetAttribute(probe_cp_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_cp_rate<0><<<grid, 32, smem>>>(reps, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, 8 * grid, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("rate SS        grid %3d: %8.1f cycles per layer GEMM of one slot (39 MMAs, N = 64)\n", grid, (double)mx / (2.0 * reps));
    CK(cudaFuncSetAttribute(probe_cp_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_cp_rate<1><<<grid, 32, smem>>>(reps, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, 8 * grid, cudaMemcpyDeviceToHost));
    mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("rate cp + TS   grid %3d: %8.1f cycles per layer GEMM of one slot (13 cp pairs shared by two slots)\n", grid, (double)mx / (2.0 * reps));
  }
  return 0;
}
