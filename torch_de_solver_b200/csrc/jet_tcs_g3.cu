// Streamed tensor-core path: instantiations of signature group 3 (see jet_tcs_kernel.cuh; split for parallel compilation).
#include "jet_tcs_kernel.cuh"

namespace tdb {

TDB_TCS_DEFINE_GROUP(launch_jet_tcs_g3, TDB_TC_SIGS_G3)

}  // namespace tdb
