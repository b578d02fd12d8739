// Streamed tensor-core path: instantiations of signature group 2 (see jet_tcs_kernel.cuh; split for parallel compilation).
#include "jet_tcs_kernel.cuh"

namespace tdb {

TDB_TCS_DEFINE_GROUP(launch_jet_tcs_g2, TDB_TC_SIGS_G2)

}  // namespace tdb
