// Shared definitions for the tedeous-b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "tdb200.h"

namespace tdb {

constexpr int kThreads = 256;          // threads per CTA of the fused jet kernel (8 warps)
constexpr int kRows = TDB200_ROWS_PER_TILE;   // (point x channel) rows per tile
constexpr int kLd = 132;               // smem row stride of activation buffers: 128 + 4 keeps 16-byte
                                       // accesses of 8 consecutive rows on distinct banks
constexpr int kWLd = 128;              // smem row stride of the weight tile: 8 warps x 16 columns
constexpr int kMaxW = 128;             // widest hidden layer
constexpr int kMaxOut = 8;
constexpr int kMaxCParams = 8;

// Everything the fused kernel needs, passed by value (fits the 4 KB kernel-parameter space).
struct JetArgs {
  int n_layers;
  int widths[TDB200_MAX_LAYERS + 1];
  int w_off[TDB200_MAX_LAYERS];        // offsets of W_l ([out][in]) and b_l in the packed parameter arena
  int b_off[TDB200_MAX_LAYERS];
  int n_net_params;                    // floats of all W, b
  int n_cparams;
  int n_params;                        // n_net_params + n_cparams
  int n_params_pad;                    // row stride of the per-CTA gradient partials
  int wmax;                            // widest hidden layer of this net
  const float* arena;                  // packed parameters, same layout as the flat gradient
  const float* arena_t;                // W_l transposed ([in][out]) at w_off[l]
  const float* img_f;                  // per layer [kMaxW][kWLd]: W^T grouped per warp (forward GEMM tile image)
  const float* img_b;                  // per layer [kMaxW][kWLd]: W grouped per warp (backward-data GEMM tile image)
  const tdb200_segment* segs;
  int n_segs;
  const int* seg_tile_begin;           // [n_segs + 1]
  int n_tiles;
  const tdb200_term* terms;
  const tdb200_factor* factors;
  int n_terms, n_factors;
  const float* comb;
  const float* pts;
  const float* targets;
  const float* coeffs;
  const float* slot_scale;             // lambda_s / len_s
  int n_slots;
  int d;
  float* part_grad;                    // [gridDim.x][n_params_pad]
  double* part_loss;                   // [gridDim.x][n_slots]
  float* scratch;                      // [gridDim.x][scratch_per_cta] saved activations (L2 resident)
  long long scratch_per_cta;
  float* fields;                       // optional per-row operator values
  const float* row_weight;             // optional per-row loss weights of segment 0 (causal loss, losses.py:137-182)
  const float* field_seed;             // optional cotangents of every field value (layout of `fields`): the gradient
                                       // becomes the vector-Jacobian product sum seed * d field / d theta
  int do_grad;
  float* jac_rows;                     // Jacobian-rows mode of the SIMT kernel: [n_groups of segment jac_seg][n_params_pad]
  int jac_seg, jac_col;
  long long* dbg;                      // optional [gridDim.x][16] phase cycle counters (TDB200_TC_TIMING=1)
};

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream still runs - once every CTA of the predecessor has executed
// pdl_launch_dependents() (or exited) and SM resources are free; everything it does before pdl_wait() must be
// independent of the predecessor's output.  pdl_wait() returns when the predecessor has completed and its memory
// operations are visible.  Used to run the set-up of jet_tcs / wgrad_gemm / reduce_partials (shared-memory zero fill,
// barrier and TMEM allocation) on SMs the predecessor has already left.  Both are no-ops in a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// launch with the programmatic-serialization attribute (TDB200_NO_PDL=1: plain launch)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  static const bool off = getenv("TDB200_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = off ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// tensor-core weight images (jet_tc_kernel.cuh): K-major SWIZZLE_128B, [4 k-blocks][104 rows][32 floats] per image
constexpr int kTcWRows = 104;                    // rows of the weight image (neurons padded to 8)
constexpr int kTcWBlock = kTcWRows * 32;
constexpr int kTcWFloats = 4 * kTcWBlock;        // 13312 floats = 52 KB
// float offset of element (row, k) inside a swizzled operand buffer with `rows` rows per k-block
__host__ __device__ __forceinline__ int sw_off(int row, int k, int rows) {
  return (k >> 5) * rows * 32 + row * 32 + ((((k & 31) >> 2) ^ (row & 7)) << 2) + (k & 3);
}

struct PackArgs {
  int n_layers;
  int widths[TDB200_MAX_LAYERS + 1];
  int w_off[TDB200_MAX_LAYERS];
  int b_off[TDB200_MAX_LAYERS];
  int n_net_params;
  int n_cparams;
  const float* W[TDB200_MAX_LAYERS];
  const float* b[TDB200_MAX_LAYERS];
  const float* c[kMaxCParams];
  float* arena;
  float* arena_t;
  float* img_f;
  float* img_b;
  float* wimg;                         // optional: hi / lo tensor-core images of the W x W layers, written by the same
                                       // launch (layout of pack_tc_images_kernel; one launch less per step)
};

// host-side launchers (jet_simt.cu)
size_t jet_simt_smem_bytes(int wmax);
cudaError_t launch_pack_params(const PackArgs& a, cudaStream_t s);
cudaError_t launch_jet_simt(const JetArgs& a, int grid, cudaStream_t s);
cudaError_t launch_reduce_partials(const float* part_grad, int n_grad_rows, const double* part_loss,
                                   int n_loss_rows, int n_params, int n_params_pad, int n_slots,
                                   const double* slot_lambda, const double* slot_len, float* out, cudaStream_t s);
// tensor-core path (jet_tc.cu)
size_t jet_tc_smem_bytes();
cudaError_t launch_pack_tc_images(const PackArgs& a, float* img, cudaStream_t s);
bool jet_tc_supports(int o0, int o1, int o2);
int jet_tc_points_per_tile(int o0, int o1, int o2);
int jet_tc_columns_per_part(int o0, int o1, int o2);   // used (point, channel) columns of the 16 a thread owns
int jet_tc_partial_rows();
int jet_tc_max_out();
cudaError_t launch_jet_tc(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s);

// tanh and its derivatives as functions of a = tanh(z):  f1 = 1 - a^2, f_{k+1} = d f_k / dz
struct TanhF {
  float f1, f2, f3, f4, f5;
  __device__ __forceinline__ explicit TanhF(float a) {
    f1 = fmaf(-a, a, 1.f);
    f2 = -2.f * a * f1;
    f3 = -2.f * fmaf(f1, f1, a * f2);
    f4 = -2.f * fmaf(3.f * f1, f2, a * f3);
    f5 = -2.f * (3.f * f2 * f2 + 4.f * f1 * f3 + a * f4);
  }
};

// Forward Taylor-mode rule through y = tanh(z) along one direction: z[0..order) = z', z'', ...
__device__ __forceinline__ void tanh_jet_fwd(const TanhF& f, const float* z, int order, float* y) {
  const float z1 = z[0];
  y[0] = f.f1 * z1;
  if (order >= 2) y[1] = f.f2 * z1 * z1 + f.f1 * z[1];
  if (order >= 3) y[2] = f.f3 * z1 * z1 * z1 + 3.f * f.f2 * z1 * z[1] + f.f1 * z[2];
  if (order >= 4)
    y[3] = f.f4 * z1 * z1 * z1 * z1 + 6.f * f.f3 * z1 * z1 * z[1] + 3.f * f.f2 * z[1] * z[1] +
           4.f * f.f2 * z1 * z[2] + f.f1 * z[3];
}

// Adjoint of tanh_jet_fwd: given gy[0..order), returns gz[0..order) and the contribution to gz0.
__device__ __forceinline__ float tanh_jet_bwd(const TanhF& f, const float* z, const float* gy, int order,
                                              float* gz) {
  const float z1 = z[0];
  float g0 = gy[0] * f.f2 * z1;
  gz[0] = gy[0] * f.f1;
  if (order >= 2) {
    const float z2 = z[1];
    g0 += gy[1] * (f.f3 * z1 * z1 + f.f2 * z2);
    gz[0] += gy[1] * 2.f * f.f2 * z1;
    gz[1] = gy[1] * f.f1;
    if (order >= 3) {
      const float z3 = z[2];
      g0 += gy[2] * (f.f4 * z1 * z1 * z1 + 3.f * f.f3 * z1 * z2 + f.f2 * z3);
      gz[0] += gy[2] * (3.f * f.f3 * z1 * z1 + 3.f * f.f2 * z2);
      gz[1] += gy[2] * 3.f * f.f2 * z1;
      gz[2] = gy[2] * f.f1;
      if (order >= 4) {
        const float z4 = z[3];
        g0 += gy[3] * (f.f5 * z1 * z1 * z1 * z1 + 6.f * f.f4 * z1 * z1 * z2 + 3.f * f.f3 * z2 * z2 +
                       4.f * f.f3 * z1 * z3 + f.f2 * z4);
        gz[0] += gy[3] * (4.f * f.f4 * z1 * z1 * z1 + 12.f * f.f3 * z1 * z2 + 4.f * f.f2 * z3);
        gz[1] += gy[3] * (6.f * f.f3 * z1 * z1 + 6.f * f.f2 * z2);
        gz[2] += gy[3] * 4.f * f.f2 * z1;
        gz[3] = gy[3] * f.f1;
      }
    }
  }
  return g0;
}

__device__ __forceinline__ float pow_i(float x, int ipow, float p) {
  if (ipow < 0) return powf(x, p);
  float r = 1.f;
  for (int i = 0; i < ipow; ++i) r *= x;
  return r;
}
__device__ __forceinline__ float dpow_i(float x, int ipow, float p) {
  if (ipow < 0) return p * powf(x, p - 1.f);
  if (ipow == 0) return 0.f;
  float r = (float)ipow;
  for (int i = 1; i < ipow; ++i) r *= x;
  return r;
}

}  // namespace tdb
