// Streamed tensor-core path ("tcs"): argument blocks shared by jet_tcs_kernel.cuh, wgrad_gemm.cu and capi.cu.
#pragma once
#include "common.cuh"

namespace tdb {

constexpr int kTcsMaxMma = TDB200_MAX_LAYERS - 2;   // W x W layers the streamed path handles
constexpr int kTcsMaxSegs = 8;                      // segments one jet_tcs_kernel launch can cover

// jet_tcs_kernel: fused forward + operator + backward-data over the tiles [tile0, tile1) of the interior segment.
struct TcsArgs {
  const float* wimg;            // weight images, layout of pack_tc_images_kernel
  float* ys;                    // Y_l of the chunk, l = 0..n_mma-1 (input of W x W layer l + 1): per layer float4
                                // [(tile * 4 + part) * Q + q][Wp neurons] = 4 consecutive (point, channel) columns
  float* gs;                    // gZ_t, t = 1..n_mma at index t - 1, same layout
  long long stream_stride;      // floats between two layers' arrays
  float* zsave;                 // per CTA [2 slots][n_mma][512 threads][16]: pre-activation jets (L2-resident scratch)
  int tile0, tile1;             // tiles of this launch (chunk)
  int Wp;                       // row pitch of the streams (W rounded up to 4)
  int zero_partials;            // 1: first chunk of a call (partial rows start from zero), 0: accumulate
  int w_in_tmem;                // set by the launcher: weights as tensor-memory operands (tcgen05.cp), see jet_tcs_kernel.cuh
  // Segments of this launch (a.segs is the plan's full segment array): segment mseg_index[m] owns the launch tiles
  // [mseg_tile_begin[m], mseg_tile_begin[m + 1]).  m = 0 sets the jet signature; the others are identity segments WITHOUT
  // derivative channels (Dirichlet / data value rows) or with the same directions, evaluated in the same tile shape - one
  // launch instead of one small launch pair per boundary segment.
  int n_msegs;
  int mseg_tile_begin[kTcsMaxSegs + 1];
  int mseg_index[kTcsMaxSegs];
  int term_end;                 // terms [0, term_end) are pre-decoded for the operator warps (all segments of the launch)
  int slot_base;                // loss slots of the launch lie in [slot_base, slot_base + TDB200_MAX_COLS)
};

// wgrad_gemm_kernel
struct WgradArgs {
  const float* gs;
  const float* ys;
  long long stream_stride;
  long long total4;             // float4 per layer array of this chunk = tiles * 4 * Q * Wp
  int kb;                       // K rows (stream columns) per pipeline stage: 24 or 32
  int W, Wp, n_mma, splits;     // CTAs: blockIdx = split * n_mma + layer
  float* part;                  // gradient partial rows of this kernel: [grid][n_params_pad]
  int n_params_pad;
  int w_off[kTcsMaxMma];        // offset of W_{t} (t = 1..n_mma at index t - 1) in the flat gradient
  int accumulate;               // 0: first chunk (overwrite), 1: add
};

size_t jet_tcs_smem_bytes();
int jet_tcs_threads();
cudaError_t launch_jet_tcs(const JetArgs& a, const TcsArgs& x, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_jet_tcs_g0(const JetArgs& a, const TcsArgs& x, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_jet_tcs_g1(const JetArgs& a, const TcsArgs& x, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_jet_tcs_g2(const JetArgs& a, const TcsArgs& x, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_jet_tcs_g3(const JetArgs& a, const TcsArgs& x, int o0, int o1, int o2, int grid, cudaStream_t s);
cudaError_t launch_wgrad_gemm(const WgradArgs& a, int grid, cudaStream_t s);
int wgrad_kb();                 // K rows (stream columns) per pipeline stage the kernel was built with

}  // namespace tdb
