// Micro-benchmark / layout probe for tcgen05.mma on sm_100a (not product code; evidence for DESIGN.md choices).
//   1. cycles per tcgen05.mma (M = 128) as a function of N, for smem-smem (SS) and tmem-smem (TS) operands,
//      kind::tf32 and kind::f16, measured with clock64 around batches of asynchronous MMAs + one commit.
//   2. does one shared-memory image in the SWIZZLE_128B_BASE32B pattern serve BOTH as a K-major operand
//      (forward GEMM) and as an MN-major operand (transposed GEMM) for kind::tf32?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
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
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
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
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
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

// ---------------------------------------------------------------------------------------------------
// 1. rate probe: one warp; fully unrolled batches of 39 asynchronous MMAs (13 K-steps x 3 passes, compile-time
// descriptor offsets) followed by one commit + wait.  out[0] = total cycles, out[1] = cycles spent issuing.
// MODE 0: SS tf32, 1: TS tf32 (A in TMEM), 2: SS bf16, 3: SS tf32 alternating between two accumulators
// ---------------------------------------------------------------------------------------------------
template <int N, int MODE>
__global__ void __launch_bounds__(32) probe_rate(int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  float* A = reinterpret_cast<float*>(base);                 // 64 KB
  float* B = reinterpret_cast<float*>(base + 65536);         // 128 KB (N up to 256, 4 k-blocks)
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 65536 + 131072);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (65536 + 131072) / 4; i += 32) A[i] = 0.f;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tptr)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  constexpr uint32_t idesc = MODE == 2 ? idesc_bf16(128, N, 0, 0) : idesc_tf32(128, N, 0, 0);
  const uint64_t da = umma_desc(smem_u32(A), 16, 1024, 2), db = umma_desc(smem_u32(B), 16, 1024, 2);
  const bool leader = elect_one();
  uint32_t phase = 0;
  long long t0 = 0, t_issue = 0;
  for (int r = -2; r < reps; ++r) {               // two warm-up batches
    if (r == 0) { t0 = clock64(); t_issue = 0; }
    const long long ta = clock64();
#pragma unroll
    for (int i = 0; i < 39; ++i) {
      const int s = i % 13;
      const uint64_t ao = ((uint64_t)(s >> 2) * 128 * 128 + (uint64_t)(s & 3) * 32) >> 4;
      const uint64_t bo = ((uint64_t)(s >> 2) * N * 128 + (uint64_t)(s & 3) * 32) >> 4;
      if (leader) {
        if (MODE == 0) mma_tf32_ss(tmem, da + ao, db + bo, idesc, i ? 1u : 0u);
        else if (MODE == 1) mma_tf32_ts(tmem, tmem + 256 + (uint32_t)s * 8, db + bo, idesc, i ? 1u : 0u);
        else if (MODE == 2) mma_f16_ss(tmem, da + ao, db + bo, idesc, i ? 1u : 0u);
        else mma_tf32_ss(tmem + (uint32_t)(i & 1) * 256, da + ao, db + bo, idesc, i > 1 ? 1u : 0u);
      }
    }
    t_issue += clock64() - ta;
    if (leader) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t_issue; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// 2. layout probe
// ---------------------------------------------------------------------------------------------------
static inline int sw_off(int row, int k, int rows) {       // K-major SWIZZLE_128B
  return (k >> 5) * rows * 32 + row * 32 + ((((k & 31) >> 2) ^ (row & 7)) << 2) + (k & 3);
}
static inline int sw_off_b32(int r, int k, int rows) {     // SWIZZLE_128B_BASE32B pattern
  return (k >> 5) * rows * 32 + r * 32 + ((((k & 31) >> 3) ^ (r & 3)) << 3) + (k & 7);
}

// test: 0 = reference (A K-major SW128 image, type 2), 1 = A BASE32B image read K-major (type 1),
//       2 = A BASE32B image read MN-major (type 1): D[k][c] = sum_n A[n][k] B[c][n]
__global__ void __launch_bounds__(128) probe_layout(const float* __restrict__ Aimg, const float* __restrict__ Bimg,
                                                    int test, int a_sbo, float* __restrict__ D) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  float* A = reinterpret_cast<float*>(base);                 // [4][128][32]
  float* B = reinterpret_cast<float*>(base + 65536);         // [4][48][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 65536 + 4 * 48 * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4 * 128 * 32; i += 128) A[i] = Aimg[i];
  for (int i = tid; i < 4 * 48 * 32; i += 128) B[i] = Bimg[i];
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(tptr)) : "memory");
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
      const uint32_t bo = (uint32_t)(s >> 2) * 48 * 128 + (uint32_t)(s & 3) * 32;
      const uint64_t db = umma_desc(smem_u32(B) + bo, 16, 1024, 2);
      uint64_t da;
      uint32_t idesc;
      if (test == 2) {
        da = umma_desc(smem_u32(A) + (uint32_t)s * 1024, 128 * 128, a_sbo, 1);
        idesc = idesc_tf32(128, 48, 1, 0);
      } else {
        const uint32_t ao = (uint32_t)(s >> 2) * 128 * 128 + (uint32_t)(s & 3) * 32;
        da = umma_desc(smem_u32(A) + ao, 16, a_sbo, test == 0 ? 2 : 1);
        idesc = idesc_tf32(128, 48, 0, 0);
      }
      if (leader) mma_tf32_ss(tmem, da, db, idesc, s ? 1u : 0u);
    }
    if (leader) umma_commit(bar);
    __syncwarp();
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 48; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * 48 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem) : "memory");
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : -1;   // -1: rate probes, 0..3: one layout test (separate processes: a bad layout faults)
  long long* d_out;
  CK(cudaMalloc(&d_out, 16 * 256));
  const size_t smem = 65536 + 131072 + 1024 + 64;
  const int reps = 50;
  if (which < 0) {
#define RUN(NN, MODE, GRID, LABEL)                                                                         \
  do {                                                                                                     \
    CK(cudaFuncSetAttribute(probe_rate<NN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    probe_rate<NN, MODE><<<GRID, 32, smem>>>(reps, d_out);                                                 \
    CK(cudaDeviceSynchronize());                                                                           \
    long long h[2 * 148];                                                                                  \
    CK(cudaMemcpy(h, d_out, 16 * GRID, cudaMemcpyDeviceToHost));                                           \
    long long mx = 0, mi = 0;                                                                              \
    for (int i = 0; i < GRID; ++i) { if (h[2 * i] > mx) { mx = h[2 * i]; mi = h[2 * i + 1]; } }            \
    printf("rate %-10s grid %3d M=128 N=%3d: %7.1f cycles / MMA total, %6.1f issuing (math floor %5.1f)\n", LABEL, GRID, \
           NN, (double)mx / (39.0 * reps), (double)mi / (39.0 * reps), NN / 2.0);                           \
  } while (0)
#define RUN_ALL(MODE, LABEL) \
  RUN(16, MODE, 1, LABEL); RUN(32, MODE, 1, LABEL); RUN(48, MODE, 1, LABEL); RUN(64, MODE, 1, LABEL); RUN(96, MODE, 1, LABEL); \
  RUN(128, MODE, 1, LABEL); RUN(192, MODE, 1, LABEL); RUN(256, MODE, 1, LABEL)
  RUN_ALL(0, "SS tf32");
  RUN_ALL(1, "TS tf32");
  RUN_ALL(2, "SS bf16");
  RUN_ALL(3, "SS tf32 x2");
  RUN(96, 0, 148, "SS tf32");
  return 0;
  }
  // ---- layout probe ----
  std::vector<float> Am(128 * 128), Bm(48 * 128);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) Am[n * 128 + k] = (float)(((n * 7 + k * 3) % 17) - 8);
  for (int c = 0; c < 48; ++c)
    for (int k = 0; k < 128; ++k) Bm[c * 128 + k] = (float)(((c * 5 + k) % 13) - 6);
  std::vector<float> imgA_sw(4 * 128 * 32), imgA_b32(4 * 128 * 32), imgB(4 * 48 * 32);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) {
      imgA_sw[sw_off(n, k, 128)] = Am[n * 128 + k];
      imgA_b32[sw_off_b32(n, k, 128)] = Am[n * 128 + k];
    }
  for (int c = 0; c < 48; ++c)
    for (int k = 0; k < 128; ++k) imgB[sw_off(c, k, 48)] = Bm[c * 128 + k];
  float *dA, *dA2, *dB, *dD;
  CK(cudaMalloc(&dA, imgA_sw.size() * 4));
  CK(cudaMalloc(&dA2, imgA_b32.size() * 4));
  CK(cudaMalloc(&dB, imgB.size() * 4));
  CK(cudaMalloc(&dD, 128 * 48 * 4));
  CK(cudaMemcpy(dA, imgA_sw.data(), imgA_sw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA2, imgA_b32.data(), imgA_b32.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, imgB.data(), imgB.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem2 = 65536 + 4 * 48 * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  struct T { int test; int sbo; const char* what; };
  const T tests[] = {{0, 1024, "K-major SWIZZLE_128B image (known-good harness check)"},
                     {1, 1024, "BASE32B image read K-major, SBO 1024"},
                     {1, 512, "BASE32B image read K-major, SBO 512"},
                     {2, 512, "BASE32B image read MN-major, SBO 512"}};
  for (int ti = 0; ti < 4; ++ti) {
    if (ti != which) continue;
    const T& t = tests[ti];
    CK(cudaMemset(dD, 0, 128 * 48 * 4));
    probe_layout<<<1, 128, smem2>>>(t.test == 0 ? dA : dA2, dB, t.test, t.sbo, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("layout %-55s: CUDA error %s\n", t.what, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * 48);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < 104; ++m)
      for (int c = 0; c < 48; ++c) {
        double ref = 0;
        for (int j = 0; j < 104; ++j)
          ref += t.test == 2 ? (double)Am[j * 128 + m] * Bm[c * 128 + j] : (double)Am[m * 128 + j] * Bm[c * 128 + j];
        const double err = fabs(ref - D[m * 48 + c]);
        if (err > 1e-3) ++bad;
        if (err > maxerr) maxerr = err;
      }
    printf("layout %-55s: %s (%d of %d wrong, max err %.3g)\n", t.what, bad ? "MISMATCH" : "exact", bad, 104 * 48, maxerr);
  }
  return 0;
}
