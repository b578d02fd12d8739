// tcgen05 path: instantiations of signature group 1 (see jet_tc_kernel.cuh; split for parallel compilation).
#include "jet_tc_kernel.cuh"

namespace tdb {

TDB_TC_DEFINE_GROUP(launch_jet_tc_g1, TDB_TC_SIGS_G1)

}  // namespace tdb
