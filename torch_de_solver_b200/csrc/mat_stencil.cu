// mat-mode stencil kernels - placeholder entry points until the kernel lands (next commit).
#include <string>
#include "common.cuh"
extern "C" {
int tdb200_mat_plan_create(const tdb200_mat_desc*, const tdb200_mat_field*, int32_t, const float*, const int32_t*,
                           const int32_t*, int32_t, const tdb200_term*, int32_t, const tdb200_factor*, int32_t,
                           tdb200_mat_plan**) { return TDB200_ERR_INVALID; }
int tdb200_mat_plan_set_coeffs(tdb200_mat_plan*, const float*, int64_t) { return TDB200_ERR_INVALID; }
int tdb200_mat_plan_set_bcs(tdb200_mat_plan*, int32_t, const tdb200_mat_bc*, const int32_t*, const float*, int32_t,
                            const double*, const double*) { return TDB200_ERR_INVALID; }
int tdb200_mat_loss_grad(tdb200_mat_plan*, const float*, float*, float*, void*) { return TDB200_ERR_INVALID; }
int64_t tdb200_mat_plan_out_size(const tdb200_mat_plan*) { return 0; }
int32_t tdb200_mat_plan_launches_per_call(const tdb200_mat_plan*) { return 0; }
void tdb200_mat_plan_destroy(tdb200_mat_plan*) {}
}
