// Peer-memory exchange for sharded mat mode (SURVEY 8e "direct peer loads over NVLink"): the halo rows of the slab
// decomposition and the handful of loss terms move between the GPUs of one box through CUDA-IPC mapped buffers, written
// and read by two tiny kernels of our own - no collective library call on the step's critical path.
//
// Every rank owns one exchange block (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by every other rank):
//     [ flags | loss inbox [2 parities][world][kPeerMaxLoss] | halo inbox [2 parities][2 sides][halo floats] ]
// PUSH model: data and flags are WRITTEN into the receiver's block (posted NVLink stores), every wait polls LOCAL memory
// (the first version pulled: remote polls and dependent remote loads cost 17.7 us per halo exchange, this one ~6).
// A step on rank r:
//   peer_halo_kernel   one CTA per side: store my first / last owned halo rows into the neighbour's inbox (parity =
//                      step & 1), system fence, release-store the step number into the neighbour's flag; wait for my
//                      own flag of that side; copy my inbox rows into my extended slab.  The neighbour cannot run two
//                      steps ahead (its next push waits for my flag), so two parities are enough.
//   ... the unchanged stencil / boundary kernels on the extended slab ...
//   peer_loss_kernel   one CTA: store my [2 + n_slots] loss terms into every rank's inbox row `rank`, fence, flags; wait
//                      for every rank's flag; add the rows in RANK order (every rank gets the bit-identical sum).
// Both kernels are plain stream work: the whole multi-rank step is capturable in one CUDA graph.  Waits are bounded
// (~20 s of polling): on a time-out the kernel sets an error word instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "tdb200.h"

extern "C" void tdb200_set_error_(const char* msg);

#include "peer_dev.cuh"

namespace tdb {
}  // namespace tdb

struct tdb200_peer {
  int rank = 0, world = 1, device = 0;
  long long halo_floats = 0;                    // floats of one side's halo block
  long long vec_floats = 0;                     // capacity of the vector all-reduce (floats, multiple of 4)
  void* mine = nullptr;                         // my exchange block
  void* blocks[tdb::kPeerMaxWorld] = {};        // every rank's block as seen from this device (mine included)
  bool opened[tdb::kPeerMaxWorld] = {};
  unsigned int* step_dev = nullptr;             // device step counters (halo, loss, halo-leave, vec, vec-arrive, vec-leave)
};

namespace tdb {

__device__ __forceinline__ float* halo_box(void* block, int parity, int side, long long halo_floats) {
  return reinterpret_cast<float*>(reinterpret_cast<char*>(block) + sizeof(PeerHeader)) + ((size_t)parity * 2 + side) * halo_floats;
}

// side 0: exchange with rank - 1 (my FIRST owned rows go out, its LAST owned rows come in above me); side 1: rank + 1.
// rows: n_var blocks of `rows_floats` contiguous floats (h rows x n1) at stride `var_stride` inside the extended slab.
__global__ void __launch_bounds__(1024) peer_halo_kernel(void* mine, void* up_block, void* down_block, unsigned int* step_dev,
                                                         float* ext, long long var_stride, int n_var, long long rows_floats,
                                                         long long own_first, long long own_last, long long halo_up,
                                                         long long halo_down, long long halo_floats) {
  // the stencil launch behind may start now: its warps that read halo rows wait for this grid (griddepcontrol.wait)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int side = blockIdx.x;
  void* const nb = side == 0 ? up_block : down_block;
  const unsigned int step = step_dev[0] + 1;
  const int parity = step & 1;
  PeerHeader* const hdr = reinterpret_cast<PeerHeader*>(mine);
  __shared__ unsigned int ok;
  if (nb) {
    // push: my rows -> the neighbour's inbox of ITS side towards me (1 - side)
    float* box = halo_box(nb, parity, 1 - side, halo_floats);
    const long long src0 = side == 0 ? own_first : own_last;
    for (int v = 0; v < n_var; ++v) {
      const float4* s4 = reinterpret_cast<const float4*>(ext + v * var_stride + src0);
      float4* d4 = reinterpret_cast<float4*>(box + v * rows_floats);
      for (long long i = threadIdx.x; i < rows_floats / 4; i += blockDim.x) d4[i] = s4[i];
    }
    // one thread fences at system scope after the block barrier (the barrier orders the other threads' stores before
    // it; a fence per thread - 1024 membar.sys - costs tens of microseconds), then the flag goes out
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      st_release_sys(&reinterpret_cast<PeerHeader*>(nb)->halo_flag[1 - side], step);
      ok = wait_seq(&hdr->halo_flag[side], step) ? 1u : 0u;               // local poll: the neighbour's rows are in
      if (!ok) hdr->error = 1;
    }
    __syncthreads();
    if (ok) {
      const float4* in4 = reinterpret_cast<const float4*>(halo_box(mine, parity, side, halo_floats));
      const long long dst0 = side == 0 ? halo_up : halo_down;
      for (int v = 0; v < n_var; ++v) {
        float4* d4 = reinterpret_cast<float4*>(ext + v * var_stride + dst0);
        for (long long i = threadIdx.x; i < rows_floats / 4; i += blockDim.x) d4[i] = __ldcv(in4 + v * (rows_floats / 4) + i);
      }
    }
  }
  // the step counter advances once both CTAs are through (the last one to leave bumps it)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* const leave = step_dev + 2;
    if (atomicAdd(leave, 1u) + 1 == 2u * step) step_dev[0] = step;
  }
}

__global__ void __launch_bounds__(128) peer_loss_kernel(PeerBlocks blocks, int rank, int world, unsigned int* step_dev,
                                                        float* out, int n) {
  peer_loss_block(blocks, rank, world, step_dev, out, n);
}

// All-reduce (sum) of a vector of n floats (the [loss terms | gradient] vector of the NN / autograd modes, 82 KB for the
// 2-100-100-100-1 net) over all ranks: every CTA pushes its slice of the vector into row `rank` of every rank's inbox,
// the last CTA to finish (device-scope counter) fences and release-stores the step number into every rank's flag; then
// every CTA waits for all local flags and adds its slice of the `world` rows in rank order (bit-identical on every
// rank).  vec inbox: [2 parities][world][vec_floats] behind the halo inbox.
__global__ void __launch_bounds__(256) peer_vec_kernel(PeerBlocks blocks, int rank, int world, unsigned int* step_dev,
                                                       float* vec, int n, long long vec_off_floats, long long vec_floats) {
  const unsigned int step = step_dev[3] + 1;
  const int parity = step & 1;
  PeerHeader* const hdr = reinterpret_cast<PeerHeader*>(blocks.b[rank]);
  const int n4 = (n + 3) / 4;
  const int per = (n4 + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n4, i0 + per);
  auto row_of = [&](void* block, int r) {
    return reinterpret_cast<float4*>(reinterpret_cast<float*>(reinterpret_cast<char*>(block) + sizeof(PeerHeader)) + vec_off_floats +
                                     ((size_t)parity * world + r) * vec_floats);
  };
  // the tail of the last float4 may reach past n: the caller's buffer is padded to a multiple of 4 floats
  const float4* v4 = reinterpret_cast<const float4*>(vec);
  for (int r = 0; r < world; ++r) {
    float4* dst = row_of(blocks.b[r], rank);
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) dst[i] = v4[i];
  }
  __shared__ unsigned int ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    ok = 1;
    __threadfence_system();
    if (atomicAdd(step_dev + 4, 1u) + 1 == gridDim.x) {                      // every CTA's slice is out (and fenced)
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(&reinterpret_cast<PeerHeader*>(blocks.b[r])->vec_flag[rank], step);
    }
    for (int r = 0; r < world; ++r)
      if (!wait_seq(&hdr->vec_flag[r], step)) { ok = 0; hdr->error = 1; }
  }
  __syncthreads();
  if (ok) {
    float4* out4 = reinterpret_cast<float4*>(vec);
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 t = __ldcv(row_of(blocks.b[rank], r) + i);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      out4[i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(step_dev + 5, 1u) + 1 == gridDim.x) {      // last CTA out: counters for the next call
    step_dev[4] = 0; step_dev[5] = 0; step_dev[3] = step;
  }
}

static int peer_fail(cudaError_t e, const char* what) {
  tdb200_set_error_((std::string(what) + ": " + cudaGetErrorString(e)).c_str());
  return TDB200_ERR_CUDA;
}

}  // namespace tdb

#define PCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return tdb::peer_fail(e_, #x); } while (0)

extern "C" {

int tdb200_peer_create(int32_t rank, int32_t world, int64_t halo_floats, int64_t vec_floats, int32_t device,
                       tdb200_peer** out) {
  if (!out || world < 1 || world > tdb::kPeerMaxWorld || rank < 0 || rank >= world || halo_floats < 0 || halo_floats % 4 ||
      vec_floats < 0 || vec_floats % 4) {
    tdb200_set_error_("tdb200_peer_create: bad argument (world <= 16, halo / vector floats multiples of 4)");
    return TDB200_ERR_INVALID;
  }
  PCU(cudaSetDevice(device));
  auto* p = new tdb200_peer();
  p->rank = rank; p->world = world; p->device = device; p->halo_floats = halo_floats; p->vec_floats = vec_floats;
  const size_t bytes = sizeof(tdb::PeerHeader) + ((size_t)4 * halo_floats + (size_t)2 * world * vec_floats) * sizeof(float);
  PCU(cudaMalloc(&p->mine, bytes));
  PCU(cudaMemset(p->mine, 0, bytes));
  PCU(cudaMalloc(&p->step_dev, 8 * sizeof(unsigned int)));
  PCU(cudaMemset(p->step_dev, 0, 8 * sizeof(unsigned int)));
  PCU(cudaDeviceSynchronize());
  p->blocks[rank] = p->mine;
  *out = p;
  return TDB200_OK;
}

int tdb200_peer_handle(tdb200_peer* p, void* handle_out_64_bytes) {
  if (!p || !handle_out_64_bytes) { tdb200_set_error_("null argument"); return TDB200_ERR_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  PCU(cudaSetDevice(p->device));
  PCU(cudaIpcGetMemHandle(&h, p->mine));
  memcpy(handle_out_64_bytes, &h, sizeof(h));
  return TDB200_OK;
}

int tdb200_peer_open(tdb200_peer* p, const void* handles_world_x_64_bytes) {
  if (!p || !handles_world_x_64_bytes) { tdb200_set_error_("null argument"); return TDB200_ERR_INVALID; }
  PCU(cudaSetDevice(p->device));
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank || p->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const char*>(handles_world_x_64_bytes) + (size_t)r * 64, sizeof(h));
    PCU(cudaIpcOpenMemHandle(&p->blocks[r], h, cudaIpcMemLazyEnablePeerAccess));
    p->opened[r] = true;
  }
  return TDB200_OK;
}

int tdb200_peer_halo(tdb200_peer* p, float* ext_dev, int64_t var_stride, int32_t n_var, int64_t rows_floats,
                     int64_t own_first, int64_t own_last, int64_t halo_up, int64_t halo_down, void* stream) {
  if (!p || !ext_dev) { tdb200_set_error_("null argument"); return TDB200_ERR_INVALID; }
  if (rows_floats % 4 || (int64_t)n_var * rows_floats != p->halo_floats) {
    tdb200_set_error_("tdb200_peer_halo: n_var * rows_floats must equal the halo size of tdb200_peer_create");
    return TDB200_ERR_INVALID;
  }
  void* up = p->rank > 0 ? p->blocks[p->rank - 1] : nullptr;
  void* down = p->rank < p->world - 1 ? p->blocks[p->rank + 1] : nullptr;
  if ((p->rank > 0 && !up) || (p->rank < p->world - 1 && !down)) {
    tdb200_set_error_("tdb200_peer_halo: tdb200_peer_open was not called");
    return TDB200_ERR_INVALID;
  }
  tdb::peer_halo_kernel<<<2, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      p->mine, up, down, p->step_dev, ext_dev, var_stride, n_var, rows_floats, own_first, own_last, halo_up, halo_down,
      p->halo_floats);
  PCU(cudaGetLastError());
  return TDB200_OK;
}

int tdb200_peer_allreduce(tdb200_peer* p, float* out_dev, int32_t n, void* stream) {
  if (!p || !out_dev || n < 1 || n > tdb::kPeerMaxLoss) { tdb200_set_error_("tdb200_peer_allreduce: 1..64 floats"); return TDB200_ERR_INVALID; }
  tdb::PeerBlocks b{};
  for (int r = 0; r < p->world; ++r) {
    if (!p->blocks[r]) { tdb200_set_error_("tdb200_peer_allreduce: tdb200_peer_open was not called"); return TDB200_ERR_INVALID; }
    b.b[r] = p->blocks[r];
  }
  tdb::peer_loss_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(b, p->rank, p->world, p->step_dev, out_dev, n);
  PCU(cudaGetLastError());
  return TDB200_OK;
}

int tdb200_peer_allreduce_vec(tdb200_peer* p, float* vec_dev, int64_t n, void* stream) {
  if (!p || !vec_dev || n < 1 || n > p->vec_floats) {
    tdb200_set_error_("tdb200_peer_allreduce_vec: 1 .. vec_floats (tdb200_peer_create) floats");
    return TDB200_ERR_INVALID;
  }
  tdb::PeerBlocks b{};
  for (int r = 0; r < p->world; ++r) {
    if (!p->blocks[r]) { tdb200_set_error_("tdb200_peer_allreduce_vec: tdb200_peer_open was not called"); return TDB200_ERR_INVALID; }
    b.b[r] = p->blocks[r];
  }
  const int n4 = (int)((n + 3) / 4);
  int grid = (n4 + 1023) / 1024;                 // >= 4 float4 per thread
  grid = grid < 1 ? 1 : grid > 32 ? 32 : grid;
  tdb::peer_vec_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(b, p->rank, p->world, p->step_dev, vec_dev, (int)n,
                                                                            4 * p->halo_floats, p->vec_floats);
  PCU(cudaGetLastError());
  return TDB200_OK;
}

// internal (mat_stencil.cu): what the finalizing block of the boundary kernel needs to exchange the loss terms itself
int tdb200_peer_export_(tdb200_peer* p, tdb::PeerLossArgs* out) {
  if (!p || !out) { tdb200_set_error_("null argument"); return TDB200_ERR_INVALID; }
  for (int r = 0; r < p->world; ++r) {
    if (!p->blocks[r]) { tdb200_set_error_("tdb200_peer_open was not called"); return TDB200_ERR_INVALID; }
    out->blocks.b[r] = p->blocks[r];
  }
  out->rank = p->rank; out->world = p->world; out->step_dev = p->step_dev;
  return TDB200_OK;
}

int tdb200_peer_error(tdb200_peer* p, int32_t* error_out) {
  if (!p || !error_out) { tdb200_set_error_("null argument"); return TDB200_ERR_INVALID; }
  tdb::PeerHeader h;
  PCU(cudaSetDevice(p->device));
  PCU(cudaMemcpy(&h, p->mine, 16, cudaMemcpyDeviceToHost));     // flags + error word
  *error_out = (int32_t)h.error;
  return TDB200_OK;
}

void tdb200_peer_destroy(tdb200_peer* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (int r = 0; r < p->world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->blocks[r]);
  cudaFree(p->mine);
  cudaFree(p->step_dev);
  delete p;
}

}  // extern "C"
