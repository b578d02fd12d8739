/* tdb200.h - C ABI of libtedeous_b200.so: the B200 (sm_100a) residual-loss hot path of TEDEouS.
 *
 * The reference (ITMO-NSS-team/torch_DE_solver v0.4.11) is pure Python and has no FFI layer; the seam every
 * caller goes through is `Solution.evaluate()` followed by `loss.backward()`:
 *     tedeous/solution.py:129-168      Solution.evaluate            -> tdb200_loss_grad / tdb200_mat_loss_grad
 *     tedeous/eval.py:223-232          Operator.operator_compute    -> tdb200_eval_fields (op part)
 *     tedeous/eval.py:433-461          Bounds.apply_bcs             -> tdb200_eval_fields (bval part)
 *     tedeous/derivative.py:30-132     Derivative_NN / _autograd    -> jet channels inside tdb200_loss_grad
 *     tedeous/derivative.py:135-323    Derivative_mat               -> tdb200_mat_loss_grad
 *     tedeous/finite_diffs.py:244-268  Finite_diffs.scheme_choose   -> host side (stencil combos in `comb`)
 *     tedeous/losses.py:84-135         Losses._default_loss         -> out[0..2+n_slots) of tdb200_loss_grad
 *     tedeous/optimizers/closure.py:60 loss.backward()              -> out[2+n_slots ..) (the flat gradient)
 *
 * Conventions: plain pointers and sizes only; every `*_dev` pointer is device memory owned by the caller
 * (torch); every entry returns 0 on success or a negative tdb200_status, with a message available from
 * tdb200_last_error(); launches are asynchronous on the `stream` argument (a cudaStream_t passed as void*).
 * A plan is not thread-safe; distinct plans are independent.  There is no CPU fallback.
 */
#ifndef TDB200_H
#define TDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDB200_MAX_LAYERS 16
#define TDB200_MAX_DIRS 4
#define TDB200_MAX_COLS 8
#define TDB200_MAX_K 32
#define TDB200_MAX_M 16
#define TDB200_MAX_J 16
#define TDB200_ROWS_PER_TILE 128

typedef enum {
  TDB200_OK = 0,
  TDB200_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  TDB200_ERR_CUDA = -2,         /* CUDA runtime error (see tdb200_last_error) */
  TDB200_ERR_NO_DEVICE = -3,    /* no sm_100 device */
  TDB200_ERR_ALLOC = -4
} tdb200_status;

/* One segment = a set of residual rows sharing operators and layout (see torch_de_solver_b200/plan.py).
 * A row is a group of K evaluation points; its channel values are V[m] = sum comb[m][k*J+c] * jet_c(x_k)
 * (identity != 0: K == 1 and V == jets). */
typedef struct {
  int64_t n_groups;
  int64_t pts_off;                       /* first row of `pts` */
  int64_t tgt_off;                       /* first float in `targets`, -1: zero targets */
  int64_t field_off;                     /* first float in the fields output */
  int32_t K, M, n_cols, n_dirs;
  int32_t dir_axis[TDB200_MAX_DIRS];     /* jet directions: input axis (-1: a general vector, see dir_vec) ... */
  int32_t dir_order[TDB200_MAX_DIRS];    /* ... and highest derivative order (1..4) */
  int32_t col_term_begin[TDB200_MAX_COLS];
  int32_t col_term_end[TDB200_MAX_COLS];
  int32_t col_slot[TDB200_MAX_COLS];     /* loss slot every residual column accumulates into */
  int32_t identity;
  int32_t comb_off;                      /* first float of this segment's [M][K*J] matrix in `comb` */
  float dir_vec[TDB200_MAX_DIRS][4];     /* direction vectors (unit vector of dir_axis for a pure partial): channel k of
                                            direction i is the k-th derivative ALONG dir_vec[i]; mixed partials such as
                                            d2u/dxdy (tedeous/derivative.py:92-97 takes any axis list) are lowered on the
                                            host to combinations of these (torch_de_solver_b200/plan.py lower_mixed) */
} tdb200_segment;

/* term = coeff * prod_f chan[f]^pow_f;  kind 0: immediate `coeff`, 1: per-row buffer coeffs[idx + row],
 * 2: trainable scalar number `idx` (appended after the network parameters) */
typedef struct {
  float coeff;
  int32_t kind;
  int64_t idx;
  int32_t fac_begin, fac_end;
} tdb200_term;

typedef struct {
  int32_t var;        /* network output */
  int32_t chan;       /* virtual channel */
  float pow;
  int32_t ipow;       /* integer power 0..16, or -1 -> powf */
} tdb200_factor;

typedef struct {
  int32_t n_layers;                          /* Linear layers, >= 2; tanh between them */
  int32_t widths[TDB200_MAX_LAYERS + 1];     /* d, w1, ..., n_out */
  int32_t n_cparams;                         /* trainable scalar coefficients */
} tdb200_net;

typedef struct tdb200_plan tdb200_plan;

/* Build a plan for modes 'NN' / 'autograd'.  Host arrays are copied. */
int tdb200_plan_create(const tdb200_net* net, int32_t n_segments, const tdb200_segment* segments,
                       int32_t n_terms, const tdb200_term* terms, int32_t n_factors,
                       const tdb200_factor* factors, int32_t n_comb, const float* comb, int32_t n_slots,
                       int32_t device, tdb200_plan** out);

/* Bind the (caller-owned, device) point / target / coefficient buffers.  pts is [n_pts, d] row-major. */
int tdb200_plan_set_points(tdb200_plan* plan, const float* pts_dev, int64_t n_pts, const float* targets_dev,
                           int64_t n_targets, const float* coeffs_dev, int64_t n_coeffs);

/* Per-slot lambda and 1/len (global row count of the slot): scale = lambda/len enters the gradient. */
int tdb200_plan_set_slots(tdb200_plan* plan, const double* slot_lambda, const double* slot_len);

/* Causal loss (tedeous/losses.py:137-182): optional per-row weights w_r of the interior operator rows (segment 0,
 * device pointer, n_groups floats, not differentiated): the operator part of the loss becomes sum_r w_r * res_r^2 *
 * lambda / len.  NULL switches the weights off. */
int tdb200_plan_set_row_weights(tdb200_plan* plan, const float* weights_dev);

/* Vector-Jacobian product mode: with seeds (device pointer, tdb200_plan_n_fields floats, the layout tdb200_eval_fields
 * writes) the gradient part of tdb200_loss_grad's output becomes sum_fields seed * d field / d theta - the backward of
 * any differentiable function of the per-point fields (weak-form loss tedeous/losses.py:184-228 and eval.py:195-221,
 * the Operator / Bounds seams of eval.py used directly).  The loss terms of that call are those of the plain loss.
 * NULL switches back to the loss gradient. */
int tdb200_plan_set_field_seeds(tdb200_plan* plan, const float* seeds_dev);

/* Multi-GPU (collocation rows shard over ranks, one all-reduce of [loss terms | gradient] per step; the
 * reference has no collectives).  tdb200_comm_unique_id fills 128 bytes (an ncclUniqueId) on one rank; the host hands them
 * to every rank by its own means (torch.distributed broadcast, MPI, a file); tdb200_comm_create makes this rank's NCCL
 * communicator (collective call: every rank, same id) - one per process is enough, plans borrow it with
 * tdb200_plan_set_comm (NULL: none).  A plan with a communicator enqueues ncclAllReduce(SUM) of its output vector on the
 * caller's stream right after the reduction kernel of tdb200_loss_grad / tdb200_eval_fields (graph-capturable).
 * NCCL is loaded with dlopen at the first of these calls; tdb200_comm_destroy is never called implicitly. */
typedef struct tdb200_comm tdb200_comm;
int tdb200_comm_unique_id(void* id_out_128_bytes);
int tdb200_comm_create(const void* unique_id_128_bytes, int32_t rank, int32_t world, int32_t device, tdb200_comm** out);
void tdb200_comm_destroy(tdb200_comm* comm);
int tdb200_plan_set_comm(tdb200_plan* plan, tdb200_comm* comm);

/* Choose the kernel implementation: 0 = auto, 1 = SIMT fp32, 2 = tcgen05 3xTF32 with the weight gradients of 1-2 W x W
 * layers in TMEM, 3 = streamed tcgen05 3xTF32 pair, any depth (2 / 3: error if the net / operator is not served). */
int tdb200_plan_set_impl(tdb200_plan* plan, int32_t impl);

/* Number of floats of the output vector of tdb200_loss_grad: 2 + n_slots + n_params. */
int64_t tdb200_plan_out_size(const tdb200_plan* plan);
int64_t tdb200_plan_n_params(const tdb200_plan* plan);
int64_t tdb200_plan_n_fields(const tdb200_plan* plan);
/* Which kernels serve the interior segment with the current impl setting: 1 = fp32 SIMT jet kernel, 2 = tcgen05 kernel
 * with the weight gradients in TMEM (1-2 W x W layers), 3 = streamed tcgen05 pair (jet_tcs_kernel + wgrad_gemm_kernel). */
int32_t tdb200_plan_kernel_path(const tdb200_plan* plan);
/* Kernel launches one tdb200_loss_grad call enqueues. */
int32_t tdb200_plan_launches_per_call(const tdb200_plan* plan);

/* One optimiser-step evaluation (replaces Solution.evaluate + loss.backward).
 * params_dev: host array of 2*n_layers + n_cparams device pointers (W0, b0, W1, b1, ..., c0, ...), W row-major
 * [out, in] as torch.nn.Linear stores it.
 * out_dev[0] = loss, out_dev[1] = loss_normalized (lambda == 1), out_dev[2 + s] = mean-square of slot s,
 * out_dev[2 + n_slots ...] = d loss / d params in the order of params_dev, flattened.
 * With several GPUs every rank gets partial sums that add up across ranks (slot_len is global). */
int tdb200_loss_grad(tdb200_plan* plan, const float* const* params_dev, float* out_dev, void* stream);

/* Forward only: per-row operator values (before subtracting targets) into fields_dev[n_fields]; also fills
 * out_dev[0 .. 2+n_slots) when out_dev != NULL. */
int tdb200_eval_fields(tdb200_plan* plan, const float* const* params_dev, float* fields_dev, float* out_dev,
                       void* stream);

/* Per-residual Jacobian rows (replaces the loop of one torch.autograd.grad per residual in
 * tedeous/optimizers/ngd.py:57-77 `gram_factory.jacobian`): row r of rows_dev[n_groups(segment)][n_params_pad] becomes
 * d field[r, col] / d params for the rows of one segment (0 = the interior operator rows, then the boundary segments in
 * plan order) and one residual column; the first tdb200_plan_n_params floats of a row are in the order of the flat
 * gradient, the pad is zero.  J v and J^T J are then dense products of this matrix.  fp32 SIMT kernel, one row per tile. */
int64_t tdb200_plan_n_params_pad(const tdb200_plan* plan);
int tdb200_jacobian_rows(tdb200_plan* plan, const float* const* params_dev, int32_t segment, int32_t col,
                         float* rows_dev, void* stream);

void tdb200_plan_destroy(tdb200_plan* plan);

/* ---- mat mode (tedeous/derivative.py:135-323): grid-stencil residual + adjoint --------------------- */

/* 1-D derivative operator D^order along one axis as a banded matrix: row i, offsets -b..b.
 * Interior rows share `interior`; the first/last n_edge rows have their own coefficients. */
typedef struct {
  int32_t var, axis, order;              /* which field, which axis, how many applications of D */
  int32_t half_width;                    /* b */
  int32_t n_edge;
  int32_t coef_off;                      /* into `band`: interior[2b+1], lo[n_edge][2b+1], hi[n_edge][2b+1] */
} tdb200_mat_field;                      /* derivative field F_q = D^order_axis u_var */

typedef struct {
  int32_t n_eq, n_var, n0, n1;           /* model is [n_var, n0, n1]; n_eq residual columns */
  int32_t n_fields;                      /* distinct derivative fields (field 0..n_var-1 = values) */
} tdb200_mat_desc;

typedef struct tdb200_mat_plan tdb200_mat_plan;

/* terms/factors as above with factor.chan = derivative-field index; term kind 1 buffers are [n0*n1]. */
int tdb200_mat_plan_create(const tdb200_mat_desc* desc, const tdb200_mat_field* fields, int32_t n_band,
                           const float* band, const int32_t* eq_term_begin, const int32_t* eq_term_end,
                           int32_t n_terms, const tdb200_term* terms, int32_t n_factors,
                           const tdb200_factor* factors, int32_t device, tdb200_mat_plan** out);
int tdb200_mat_plan_set_coeffs(tdb200_mat_plan* plan, const float* coeffs_dev, int64_t n_coeffs);

/* Boundary rows: row r = sum_k sign[k] * (bop or value)(cell[r*K + k]); `cells` are flat indices i0*n1+i1.
 * bc_term_begin == bc_term_end: value of `var`.  slot is relative to the boundary slots. */
typedef struct {
  int64_t n_rows;
  int64_t cell_off;                      /* into cells_dev (n_rows*K ints) */
  int64_t tgt_off;                       /* into targets_dev (n_rows floats) */
  int32_t K, var, slot;
  int32_t term_begin, term_end;
  float sign[4];
} tdb200_mat_bc;

int tdb200_mat_plan_set_bcs(tdb200_mat_plan* plan, int32_t n_bcs, const tdb200_mat_bc* bcs,
                            const int32_t* cells_dev, const float* targets_dev, int32_t n_slots,
                            const double* slot_lambda, const double* slot_len);

/* Slab decomposition over several GPUs: the plan is built on this rank's rows plus 2 * reach halo rows
 * on each interior side; only rows [row_lo, row_hi) of that extended slab enter the loss (slot_len is the global
 * count, so the per-rank outputs add up), gradients of the halo rows are not meaningful.  Default: all rows. */
int tdb200_mat_plan_set_row_window(tdb200_mat_plan* plan, int32_t row_lo, int32_t row_hi);

/* out_dev[0] = loss, [1] = loss_normalized, [2 + s] = slot mean squares (n_eq + n_bc_slots);
 * grad_dev [n_var, n0, n1] = d loss / d u. */
int tdb200_mat_loss_grad(tdb200_mat_plan* plan, const float* u_dev, float* grad_dev, float* out_dev,
                         void* stream);
/* Forward only: op_dev [n0*n1, n_eq] residual fields, bval_dev [total boundary rows] row values (either may be
 * NULL); out_dev as above. */
int tdb200_mat_eval_fields(tdb200_mat_plan* plan, const float* u_dev, float* op_dev, float* bval_dev,
                           float* out_dev, void* stream);
int64_t tdb200_mat_plan_out_size(const tdb200_mat_plan* plan);
int32_t tdb200_mat_plan_launches_per_call(const tdb200_mat_plan* plan);
/* Which residual kernel tdb200_mat_loss_grad launches: 0 = generic tiled kernel (any operator), 1 = register-tap
 * kernel (one linear constant-coefficient equation), 2 = vectorised cross-stencil kernel (1 + n1 % 4 == 0),
 * 3 = persistent TMA-pipelined cross-stencil kernel (2 + at most one forcing buffer), 4 = register-marching
 * cross-stencil kernel (3 + stencil reach <= 2; the edge frame is completed by the boundary launch, or by a small
 * launch of its own when some boundary row is not a value row on a frame cell). */
int32_t tdb200_mat_plan_kernel_kind(const tdb200_mat_plan* plan);
/* Measurement aid (bench.py roofline): with timing on, every eager tdb200_mat_loss_grad / tdb200_mat_eval_fields call
 * records CUDA events on its stream around the launch of the residual (stencil) kernel alone; tdb200_mat_plan_stencil_ms
 * waits for the last pair and returns the elapsed milliseconds.  Leave it off when capturing a CUDA graph. */
int tdb200_mat_plan_set_timing(tdb200_mat_plan* plan, int32_t on);
int tdb200_mat_plan_stencil_ms(tdb200_mat_plan* plan, float* ms_out);
/* Measurement aid: `iters` back-to-back launches of the residual (stencil) kernel(s) alone - no boundary rows, no
 * finalize - between two CUDA events on `stream`; *ms_out = mean milliseconds per launch (synchronises).  The working
 * set of BASELINE config 4 (u + forcing + gradient = 192 MiB) exceeds the 126 MB L2, so consecutive launches do not
 * feed each other from cache. */
int tdb200_mat_time_stencil(tdb200_mat_plan* plan, const float* u_dev, float* grad_dev, int32_t iters, float* ms_out,
                            void* stream);
void tdb200_mat_plan_destroy(tdb200_mat_plan* plan);

/* ---- peer-memory exchange for sharded mat mode ("direct peer loads over NVLink") ----------------------
 * One process per GPU on one box.  Every rank creates a peer object (an exchange block in its own memory), publishes its
 * 64-byte CUDA-IPC handle by the host's own means, and opens the handles of all ranks (in rank order, world x 64 bytes).
 * (halo_floats: one side's halo block of sharded mat mode, 0 if unused; vec_floats: capacity of tdb200_peer_allreduce_vec.)
 * tdb200_peer_halo then moves the halo rows of the slab decomposition (the first / last `rows_floats` owned floats of each
 * of the n_var fields, offsets in floats from ext_dev) into the neighbours' extended slabs, and tdb200_peer_allreduce
 * sums a short vector (<= 64 floats: the loss terms) over all ranks in rank order - each ONE small kernel of this library
 * (flags with system-scope release / acquire over NVLink, bounded waits), plain stream work that a CUDA graph captures.
 * They are collective: every rank must enqueue the same sequence of calls.  The reference has no multi-GPU path. */
typedef struct tdb200_peer tdb200_peer;
int tdb200_peer_create(int32_t rank, int32_t world, int64_t halo_floats, int64_t vec_floats, int32_t device,
                       tdb200_peer** out);
int tdb200_peer_handle(tdb200_peer* peer, void* handle_out_64_bytes);
int tdb200_peer_open(tdb200_peer* peer, const void* handles_world_x_64_bytes);
int tdb200_peer_halo(tdb200_peer* peer, float* ext_dev, int64_t var_stride, int32_t n_var, int64_t rows_floats,
                     int64_t own_first, int64_t own_last, int64_t halo_up, int64_t halo_down, void* stream);
int tdb200_peer_allreduce(tdb200_peer* peer, float* out_dev, int32_t n, void* stream);
/* Lets tdb200_mat_loss_grad sum its loss terms over the ranks itself: the finalizing block of the boundary kernel does
 * what tdb200_peer_allreduce does, one launch less per step.  *inline_out = 1 if the plan's kernel schedule supports it
 * (then the caller must NOT call tdb200_peer_allreduce on the same output), 0 otherwise. */
int tdb200_mat_plan_set_peer(tdb200_mat_plan* plan, tdb200_peer* peer, int32_t* inline_out);
/* Sum of a longer vector (<= vec_floats of tdb200_peer_create; the buffer must be readable / writable up to the next
 * multiple of 4 floats) over all ranks, in rank order; tdb200_plan_set_peer makes tdb200_loss_grad end with it instead of
 * ncclAllReduce (NN / autograd modes on one box). */
int tdb200_peer_allreduce_vec(tdb200_peer* peer, float* vec_dev, int64_t n, void* stream);
int tdb200_plan_set_peer(tdb200_plan* plan, tdb200_peer* peer);
int tdb200_peer_error(tdb200_peer* peer, int32_t* error_out);     /* 1: a wait timed out (a rank fell out of step) */
void tdb200_peer_destroy(tdb200_peer* peer);

/* ---- fused optimiser step --------------------------------------------------------------------------
 * Replaces optimizer.step() of torch.optim.Adam / AdamW / SGD as tedeous/optimizers/optimizer.py:44-61 configures them
 * (tedeous/model.py:174-191 drives it): one launch updates all `n_tensors` parameter tensors (device pointers, sizes in
 * floats) in place from the flat gradient `grad_dev` (the gradient part of tdb200_loss_grad's output vector, or the
 * gradient of tdb200_mat_loss_grad), plus a one-thread launch that advances the step counter - a fixed sequence that can
 * follow tdb200_loss_grad inside one CUDA graph.  kind: 0 Adam (weight decay added to the gradient), 1 AdamW (decoupled),
 * 2 SGD with momentum.  hyper_dev = [lr, beta1 (SGD: momentum), beta2, eps, weight_decay] and step_dev (steps taken so
 * far) live in device memory so that schedulers can change lr between replays; m_dev / v_dev: n floats each (SGD: m_dev
 * is the momentum buffer, v_dev unused).  Same arithmetic as torch's single-tensor implementation. */
#define TDB200_OPT_MAX_TENSORS 40
int tdb200_optimizer_step(int32_t kind, int32_t n_tensors, float* const* params_dev, const int64_t* sizes,
                          const float* grad_dev, float* m_dev, float* v_dev, const float* hyper_dev, int32_t* step_dev,
                          void* stream);

const char* tdb200_last_error(void);
int tdb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TDB200_H */
