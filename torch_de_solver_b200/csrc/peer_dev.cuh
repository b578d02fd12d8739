// Device side of the peer-memory exchange shared by csrc/peer.cu (its own kernels) and csrc/mat_stencil.cu (the
// finalizing block of the boundary kernel exchanges the loss terms itself: one launch less per step).
#pragma once
#include <cuda_runtime.h>

namespace tdb {

constexpr int kPeerMaxLoss = 64;
constexpr int kPeerMaxWorld = 16;

struct PeerHeader {
  unsigned int halo_flag[2];  // [side]: last step whose rows the neighbour on that side has pushed into my inbox
  unsigned int error;         // a wait timed out
  unsigned int pad;
  unsigned int loss_flag[kPeerMaxWorld];                 // [rank]: last step whose loss terms that rank has pushed
  unsigned int vec_flag[kPeerMaxWorld];                  // [rank]: last step whose vector that rank has pushed
  float loss[2][kPeerMaxWorld][kPeerMaxLoss];            // [parity][rank][term]
};

struct PeerBlocks { void* b[kPeerMaxWorld]; };

// what a kernel of another translation unit needs to run peer_loss_block (tdb200_peer_export_)
struct PeerLossArgs {
  PeerBlocks blocks;
  int rank, world;              // world <= 1: no exchange
  unsigned int* step_dev;       // device step counters of the peer object ([1]: loss exchange)
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// wait until *flag >= want (sequence numbers only grow); false on time-out.  The bound (~20 s of SM clocks) is far above
// any start-up skew between the ranks of one job - a rank that arrives late must find its neighbours still waiting, as
// it would inside an NCCL collective - and still finite, so a rank that died cannot hang the others' GPUs for good.
__device__ __forceinline__ bool wait_seq(const unsigned int* flag, unsigned int want) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 40000000000LL) return false;
  }
  return true;
}

// All-reduce (sum, rank order) of out[0 .. n) over the ranks, executed by ONE block of >= max(n, world) threads; every
// thread of the block must call it, `out` must be complete and visible to the block (a barrier before the call).
__device__ __forceinline__ void peer_loss_block(const PeerBlocks& blocks, int rank, int world, unsigned int* step_dev,
                                                float* out, int n) {
  __shared__ unsigned int peer_ok;
  const unsigned int step = step_dev[1] + 1;
  const int parity = step & 1;
  PeerHeader* const hdr = reinterpret_cast<PeerHeader*>(blocks.b[rank]);
  for (int i = threadIdx.x; i < world * n; i += blockDim.x) {               // my terms -> row `rank` of every inbox
    const int r = i / n, t = i - r * n;
    reinterpret_cast<PeerHeader*>(blocks.b[r])->loss[parity][rank][t] = out[t];
  }
  if (threadIdx.x == 0) peer_ok = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < world; ++r)
      if (r != rank) st_release_sys(&reinterpret_cast<PeerHeader*>(blocks.b[r])->loss_flag[rank], step);
  }
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank)
    if (!wait_seq(&hdr->loss_flag[threadIdx.x], step)) { peer_ok = 0; hdr->error = 1; }
  __syncthreads();
  if ((int)threadIdx.x < n && peer_ok) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcv(&hdr->loss[parity][r][threadIdx.x]);     // rank order: bit-identical sums
    out[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) step_dev[1] = step;
}

}  // namespace tdb
