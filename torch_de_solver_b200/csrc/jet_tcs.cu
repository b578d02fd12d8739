// Streamed tensor-core path, host side: signature dispatch of jet_tcs_kernel (kernel: jet_tcs_kernel.cuh).
#include "jet_tcs_kernel.cuh"

namespace tdb {

size_t jet_tcs_smem_bytes() { return kTsSmemBytes; }
int jet_tcs_threads() { return kTsThreads; }

TDB_TCS_DEFINE_GROUP(launch_jet_tcs_g0, TDB_TC_SIGS_G0)

cudaError_t launch_jet_tcs(const JetArgs& a, const TcsArgs& x, int o0, int o1, int o2, int grid, cudaStream_t s) {
  cudaError_t e = launch_jet_tcs_g0(a, x, o0, o1, o2, grid, s);
  if (e == cudaErrorInvalidValue) e = launch_jet_tcs_g1(a, x, o0, o1, o2, grid, s);
  if (e == cudaErrorInvalidValue) e = launch_jet_tcs_g2(a, x, o0, o1, o2, grid, s);
  if (e == cudaErrorInvalidValue) e = launch_jet_tcs_g3(a, x, o0, o1, o2, grid, s);
  return e;
}

}  // namespace tdb
