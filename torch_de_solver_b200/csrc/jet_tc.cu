// tcgen05 path, host side: weight-image packing, signature dispatch (kernel: jet_tc_kernel.cuh).
#include "jet_tc_kernel.cuh"

namespace tdb {

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
  pack_tc_images_kernel<<<148, 256, 0, s>>>(a, img);
  return cudaGetLastError();
}

size_t jet_tc_smem_bytes() { return kTcSmemBytes; }

bool jet_tc_supports(int o0, int o1, int o2) {
#define X(A, B, Cc) if (o0 == A && o1 == B && o2 == Cc) return true;
  TDB_TC_SIGS(X)
#undef X
  return false;
}
int jet_tc_points_per_tile(int o0, int o1, int o2) {
  const int ph = kTcPC / (1 + o0 + o1 + o2);
  return kTcParts * (ph > kTcMaxPts / kTcParts ? kTcMaxPts / kTcParts : ph);
}
int jet_tc_columns_per_part(int o0, int o1, int o2) {
  return jet_tc_points_per_tile(o0, o1, o2) / kTcParts * (1 + o0 + o1 + o2);
}
int jet_tc_partial_rows() { return 1; }
int jet_tc_max_out() { return kTcMaxOut; }

TDB_TC_DEFINE_GROUP(launch_jet_tc_g0, TDB_TC_SIGS_G0)

cudaError_t launch_jet_tc(const JetArgs& a, const float* wimg, int o0, int o1, int o2, int grid, cudaStream_t s) {
  cudaError_t e = launch_jet_tc_g0(a, wimg, o0, o1, o2, grid, s);
  if (e == cudaErrorInvalidValue) e = launch_jet_tc_g1(a, wimg, o0, o1, o2, grid, s);
  if (e == cudaErrorInvalidValue) e = launch_jet_tc_g2(a, wimg, o0, o1, o2, grid, s);
  if (e == cudaErrorInvalidValue) e = launch_jet_tc_g3(a, wimg, o0, o1, o2, grid, s);
  return e;
}

}  // namespace tdb
