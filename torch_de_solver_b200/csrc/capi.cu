// C ABI of libtedeous_b200.so - plan management and launch sequencing for the NN / autograd path.
// See include/tdb200.h for the contract and the reference interfaces every entry replaces.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "common.cuh"
#include "jet_tcs.cuh"

namespace tdb {
int comm_create(const void* unique_id, int rank, int world, void** comm_out);
int comm_all_reduce_sum(void* comm, float* buf, size_t n, cudaStream_t s);
void comm_destroy(void* comm);
}  // namespace tdb

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return TDB200_ERR_CUDA;
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

template <class T>
int upload(T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(T));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if (src) {
    e = cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  }
  return 0;
}
}  // namespace

struct tdb200_plan {
  tdb200_net net{};
  int device = 0;
  int n_sms = 0;
  int n_slots = 0;
  int impl = 0;
  std::vector<tdb200_segment> segs;
  std::vector<int> seg_tile_begin;
  int n_tiles = 0;
  int64_t n_fields = 0;
  // device copies of the program
  tdb200_segment* d_segs = nullptr;
  int* d_seg_tile_begin = nullptr;
  tdb200_term* d_terms = nullptr;
  tdb200_factor* d_factors = nullptr;
  float* d_comb = nullptr;
  float* d_slot_scale = nullptr;
  double* d_slot_lambda = nullptr;
  double* d_slot_len = nullptr;
  // caller-owned buffers
  const float* pts = nullptr;
  const float* targets = nullptr;
  const float* coeffs = nullptr;
  int64_t n_pts = 0;
  // workspace
  float* arena = nullptr;
  float* arena_t = nullptr;
  float* img_f = nullptr;
  float* img_b = nullptr;
  float* part_grad = nullptr;
  double* part_loss = nullptr;
  float* scratch = nullptr;
  int grid = 0;                    // SIMT kernel CTAs when it runs every segment
  tdb::JetArgs args{};
  // tensor-core path: segment 0 (interior) on tcgen05, the remaining segments on the SIMT kernel
  bool tc_eligible = false;
  int tc_tiles = 0, tc_grid = 0;
  int tc_sig[3] = {0, 0, 0};
  int simt_rest_tiles = 0, simt_rest_grid = 0;
  struct TcExtra { int seg, sig[3], tiles, grid; };      // boundary segments that also run on tcgen05 (identity rows)
  std::vector<TcExtra> tc_extra;
  cudaStream_t side = nullptr;             // boundary-row launch running next to the tcgen05 launch (fork / join by events)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int* d_seg_tile_begin_tc = nullptr;
  int* d_seg_tile_begin_rest = nullptr;
  float* wimg = nullptr;
  float* tc_scratch = nullptr;
  long long tc_scratch_per_cta = 0;
  int grad_rows = 0, loss_rows = 0;   // allocated rows of the partial buffers
  // streamed tensor-core path (jet_tcs_kernel + wgrad_gemm_kernel): any number of W x W layers
  bool tcs_eligible = false;
  std::vector<TcExtra> tcs_extra;              // boundary segments made of identity rows: streamed kernels too
  // ... unless they ride in the interior launch (value rows: no derivative channels - Dirichlet / data conditions):
  std::vector<int> tcs_mseg;                   // segments of the merged launch, interior first
  std::vector<int> tcs_mseg_tile_begin;        // their first tiles (+ total)
  int tcs_tiles = 0, tcs_term_end = 0, tcs_slot_base = 0;
  int* d_seg_tile_begin_rest_all = nullptr;    // the remaining segments (periodic / finite-difference groups): SIMT kernel
  int simt_rest_all_tiles = 0;
  float* tcs_ys = nullptr;                     // streamed Y_l / gZ_t rows of one chunk
  float* tcs_gs = nullptr;
  float* tcs_zsave = nullptr;
  long long tcs_stream_stride = 0;             // floats per layer array
  int tcs_chunk_tiles = 0;
  void* comm = nullptr;                        // NCCL communicator of the plan (tdb200_plan_comm_init), or none
  tdb200_peer* peer = nullptr;                 // peer-memory exchange (tdb200_plan_set_peer): replaces the NCCL all-reduce
};

static int points_per_tile(int J, int K) {
  int step = K;
  if (step % 4) step = (step % 2) ? step * 4 : step * 2;
  return ((tdb::kRows / J) / step) * step;
}

extern "C" {

const char* tdb200_last_error(void) { return g_err.c_str(); }
void tdb200_set_error_(const char* msg) { g_err = msg ? msg : ""; }
int tdb200_version(void) { return 100; }

int tdb200_plan_create(const tdb200_net* net, int32_t n_segments, const tdb200_segment* segments,
                       int32_t n_terms, const tdb200_term* terms, int32_t n_factors,
                       const tdb200_factor* factors, int32_t n_comb, const float* comb, int32_t n_slots,
                       int32_t device, tdb200_plan** out) {
  if (!net || !segments || !out || n_segments <= 0) return fail(TDB200_ERR_INVALID, "null argument");
  if (net->n_layers < 2 || net->n_layers > TDB200_MAX_LAYERS) return fail(TDB200_ERR_INVALID, "n_layers out of range");
  if (n_slots <= 0 || n_slots > 32) return fail(TDB200_ERR_INVALID, "n_slots must be 1..32");
  if (net->n_cparams < 0 || net->n_cparams > tdb::kMaxCParams) return fail(TDB200_ERR_INVALID, "too many coefficient parameters");
  const int L = net->n_layers;
  if (net->widths[0] < 1 || net->widths[0] > 4) return fail(TDB200_ERR_INVALID, "input dimension must be 1..4");
  if (net->widths[L] < 1 || net->widths[L] > tdb::kMaxOut) return fail(TDB200_ERR_INVALID, "output dimension must be 1..8");
  int wmax = 0;
  for (int l = 1; l < L; ++l) {
    if (net->widths[l] < 1 || net->widths[l] > tdb::kMaxW) return fail(TDB200_ERR_INVALID, "hidden width must be 1..128");
    wmax = net->widths[l] > wmax ? net->widths[l] : wmax;
  }
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) return fail(TDB200_ERR_NO_DEVICE, "no CUDA device");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(TDB200_ERR_NO_DEVICE, "tedeous-b200 kernels are built for sm_100a only");

  auto* p = new tdb200_plan();
  p->net = *net;
  p->device = device;
  p->n_sms = prop.multiProcessorCount;
  p->n_slots = n_slots;
  p->segs.assign(segments, segments + n_segments);
  p->seg_tile_begin.resize(n_segments + 1);
  int tiles = 0;
  for (int s = 0; s < n_segments; ++s) {
    const tdb200_segment& sg = segments[s];
    int J = 1;
    if (sg.n_dirs < 0 || sg.n_dirs > TDB200_MAX_DIRS) { delete p; return fail(TDB200_ERR_INVALID, "n_dirs out of range"); }
    for (int i = 0; i < sg.n_dirs; ++i) {
      if (sg.dir_order[i] < 1 || sg.dir_order[i] > 4 || sg.dir_axis[i] < -1 || sg.dir_axis[i] >= net->widths[0]) {
        delete p; return fail(TDB200_ERR_INVALID, "bad jet direction");
      }
      J += sg.dir_order[i];
    }
    if (J > TDB200_MAX_J || sg.K < 1 || sg.K > TDB200_MAX_K || sg.M < 1 || sg.M > TDB200_MAX_M ||
        sg.n_cols < 1 || sg.n_cols > TDB200_MAX_COLS || sg.M > J * sg.K) {
      delete p; return fail(TDB200_ERR_INVALID, "segment shape out of range");
    }
    if (sg.identity && (sg.K != 1 || sg.M != J)) { delete p; return fail(TDB200_ERR_INVALID, "identity segment needs K == 1, M == J"); }
    const int P = points_per_tile(J, sg.K);
    if (P <= 0) { delete p; return fail(TDB200_ERR_INVALID, "group does not fit a tile"); }
    const int G = P / sg.K;
    p->seg_tile_begin[s] = tiles;
    tiles += (int)((sg.n_groups + G - 1) / G);
    p->n_fields += sg.n_groups * sg.n_cols;
    for (int c = 0; c < sg.n_cols; ++c)
      if (sg.col_slot[c] < 0 || sg.col_slot[c] >= n_slots) { delete p; return fail(TDB200_ERR_INVALID, "slot out of range"); }
  }
  p->seg_tile_begin[n_segments] = tiles;
  p->n_tiles = tiles;

  int rc;
  if ((rc = upload(&p->d_segs, segments, n_segments))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload(&p->d_seg_tile_begin, p->seg_tile_begin.data(), p->seg_tile_begin.size()))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload(&p->d_terms, terms, n_terms))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload(&p->d_factors, factors, n_factors))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload(&p->d_comb, comb, n_comb))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload<float>(&p->d_slot_scale, nullptr, n_slots))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload<double>(&p->d_slot_lambda, nullptr, n_slots))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload<double>(&p->d_slot_len, nullptr, n_slots))) { tdb200_plan_destroy(p); return rc; }

  tdb::JetArgs& a = p->args;
  a.n_layers = L;
  int off = 0;
  for (int l = 0; l <= L; ++l) a.widths[l] = net->widths[l];
  for (int l = 0; l < L; ++l) {
    a.w_off[l] = off; off += net->widths[l] * net->widths[l + 1];
    a.b_off[l] = off; off += net->widths[l + 1];
  }
  a.n_net_params = off;
  a.n_cparams = net->n_cparams;
  a.n_params = off + net->n_cparams;
  a.n_params_pad = (a.n_params + 31) / 32 * 32;
  a.wmax = wmax;
  a.n_segs = n_segments;
  a.n_tiles = tiles;
  a.n_slots = n_slots;
  a.d = net->widths[0];
  a.scratch_per_cta = (long long)2 * (L - 1) * wmax * tdb::kRows;

  p->grid = tiles < p->n_sms ? (tiles > 0 ? tiles : 1) : p->n_sms;
  {  // can the interior segment run on the tensor cores?
    const tdb200_segment& s0 = segments[0];
    bool ok = L >= 3 && net->widths[1] <= 104 && net->widths[L] <= tdb::jet_tc_max_out() && net->widths[0] <= 4;
    for (int l = 2; l < L; ++l) ok = ok && net->widths[l] == net->widths[1];
    for (int i = 0; i < 3; ++i) p->tc_sig[i] = i < s0.n_dirs ? s0.dir_order[i] : 0;
    ok = ok && s0.identity && s0.K == 1 && s0.n_dirs <= 3 && s0.n_groups > 0 &&
         tdb::jet_tc_supports(p->tc_sig[0], p->tc_sig[1], p->tc_sig[2]) &&
         s0.col_term_end[s0.n_cols - 1] <= 48 && n_terms >= 0;
    if (ok)   // every factor of the interior segment must sit in the first 96 entries (cached in shared memory)
      for (int t = 0; t < s0.col_term_end[s0.n_cols - 1]; ++t) ok = ok && terms[t].fac_end <= 96;
    p->tcs_eligible = ok;                      // streamed path: any depth
    p->tc_eligible = ok && L - 2 <= 2;         // dW accumulators of <= 2 W x W layers fit TMEM next to D and gZ
    if (ok) {
      const int P = tdb::jet_tc_points_per_tile(p->tc_sig[0], p->tc_sig[1], p->tc_sig[2]);
      p->tc_tiles = (int)((s0.n_groups + P - 1) / P);
      p->tc_grid = p->tc_tiles < p->n_sms ? p->tc_tiles : p->n_sms;
      if (getenv("TDB200_TC_GRID")) {                       // experiment knob: fewer persistent CTAs (contention studies)
        const int g = atoi(getenv("TDB200_TC_GRID"));
        if (g >= 1 && g < p->tc_grid) p->tc_grid = g;
      }
      std::vector<int> tb(n_segments + 1, p->tc_tiles), rb(n_segments + 1, 0), ra(n_segments + 1, 0);
      tb[0] = 0;
      p->tcs_mseg.assign(1, 0);
      p->tcs_mseg_tile_begin.assign(1, 0);
      p->tcs_tiles = p->tc_tiles;
      p->tcs_term_end = s0.col_term_end[s0.n_cols - 1];
      int slot_lo = s0.col_slot[0], slot_hi = s0.col_slot[0];
      for (int c = 0; c < s0.n_cols; ++c) { slot_lo = std::min(slot_lo, s0.col_slot[c]); slot_hi = std::max(slot_hi, s0.col_slot[c]); }
      for (int s = 1; s < n_segments; ++s) {
        const tdb200_segment& sg = segments[s];
        int sig[3];
        for (int i = 0; i < 3; ++i) sig[i] = i < sg.n_dirs ? sg.dir_order[i] : 0;
        bool eok = sg.identity && sg.K == 1 && sg.n_dirs <= 3 && sg.n_groups > 0 && sg.n_cols >= 1 &&
                   tdb::jet_tc_supports(sig[0], sig[1], sig[2]) && sg.col_term_end[sg.n_cols - 1] <= 48 &&
                   !getenv("TDB200_NO_TC_BOUNDARY");
        if (eok)
          for (int t = sg.col_term_begin[0]; t < sg.col_term_end[sg.n_cols - 1]; ++t) eok = eok && terms[t].fac_end <= 96;
        // value rows join the interior launch (same tile shape, their derivative channels are simply unused)
        // ... when they are few: a value row costs a full jet tile slot there (Navier-Stokes at 10^6 points has 6 10^4
        // boundary rows - 4x fewer tiles and no jet arithmetic in a launch of their own signature (0, 0, 0))
        const long long merged_tiles = (sg.n_groups + P - 1) / P;
        const bool few = (p->tcs_tiles - p->tc_tiles) + merged_tiles <= std::max<long long>(p->n_sms, p->tc_tiles / 128);
        if (eok && sg.n_dirs == 0 && few && (int)p->tcs_mseg.size() < tdb::kTcsMaxSegs && !getenv("TDB200_TCS_NO_MERGE")) {
          int lo = slot_lo, hi = slot_hi;
          for (int c = 0; c < sg.n_cols; ++c) { lo = std::min(lo, sg.col_slot[c]); hi = std::max(hi, sg.col_slot[c]); }
          if (hi - lo < TDB200_MAX_COLS) {
            slot_lo = lo; slot_hi = hi;
            p->tcs_mseg.push_back(s);
            p->tcs_mseg_tile_begin.push_back(p->tcs_tiles);
            p->tcs_tiles += (int)((sg.n_groups + P - 1) / P);
            p->tcs_term_end = std::max(p->tcs_term_end, sg.col_term_end[sg.n_cols - 1]);
            ra[s + 1] = ra[s];
            continue;
          }
        }
        if (eok) {
          const int Pe = tdb::jet_tc_points_per_tile(sig[0], sig[1], sig[2]);
          tdb200_plan::TcExtra e{};
          e.seg = s; e.sig[0] = sig[0]; e.sig[1] = sig[1]; e.sig[2] = sig[2];
          e.tiles = (int)((sg.n_groups + Pe - 1) / Pe);
          e.grid = e.tiles < p->n_sms ? e.tiles : p->n_sms;
          p->tcs_extra.push_back(e);
        }
        ra[s + 1] = ra[s] + (eok ? 0 : p->seg_tile_begin[s + 1] - p->seg_tile_begin[s]);
      }
      p->tcs_mseg_tile_begin.push_back(p->tcs_tiles);
      p->tcs_slot_base = slot_lo;
      p->simt_rest_all_tiles = ra[n_segments];
      if ((rc = upload(&p->d_seg_tile_begin_rest_all, ra.data(), ra.size()))) { tdb200_plan_destroy(p); return rc; }
      // boundary segments made of identity rows (Dirichlet values, autograd-mode operator conditions) take the tcgen05
      // kernel too - one small launch each on the side stream; periodic / finite-difference groups stay on the SIMT kernel
      // ... when the interior launch is short (< ~1 ms): there the ~170 us SIMT boundary tile is the critical path.  Next
      // to a long interior launch one SIMT CTA hides all boundary rows, and the tcgen05 side launches measured 4 % slower
      // (wave, 10^6 points: 8.25 -> 8.57 ms).
      const double tc_us_plan = 25.0 + 15.0 * ((p->tc_tiles + p->tc_grid - 1) / p->tc_grid);
      const bool tc_boundary = p->tc_eligible && tc_us_plan < 1000.0 && !getenv("TDB200_NO_TC_BOUNDARY");
      for (int s = 1; s < n_segments; ++s) {
        const tdb200_segment& sg = segments[s];
        int sig[3];
        for (int i = 0; i < 3; ++i) sig[i] = i < sg.n_dirs ? sg.dir_order[i] : 0;
        bool eok = sg.identity && sg.K == 1 && sg.n_dirs <= 3 && sg.n_groups > 0 && sg.n_cols >= 1 &&
                   tdb::jet_tc_supports(sig[0], sig[1], sig[2]) && sg.col_term_end[sg.n_cols - 1] <= 48 && tc_boundary;
        if (eok)
          for (int t = sg.col_term_begin[0]; t < sg.col_term_end[sg.n_cols - 1]; ++t) eok = eok && terms[t].fac_end <= 96;
        const int simt_tiles = p->seg_tile_begin[s + 1] - p->seg_tile_begin[s];
        if (eok) {
          const int Pe = tdb::jet_tc_points_per_tile(sig[0], sig[1], sig[2]);
          tdb200_plan::TcExtra e{};
          e.seg = s; e.sig[0] = sig[0]; e.sig[1] = sig[1]; e.sig[2] = sig[2];
          e.tiles = (int)((sg.n_groups + Pe - 1) / Pe);
          e.grid = e.tiles < 16 ? e.tiles : 16;                  // most CTAs a call gives it (partial rows are sized for it)
          p->tc_extra.push_back(e);
        }
        rb[s + 1] = rb[s] + (eok ? 0 : simt_tiles);
      }
      p->simt_rest_tiles = rb[n_segments];
      p->simt_rest_grid = p->simt_rest_tiles < p->n_sms ? p->simt_rest_tiles : p->n_sms;
      if ((rc = upload(&p->d_seg_tile_begin_tc, tb.data(), tb.size()))) { tdb200_plan_destroy(p); return rc; }
      if ((rc = upload(&p->d_seg_tile_begin_rest, rb.data(), rb.size()))) { tdb200_plan_destroy(p); return rc; }
      const size_t wimg_floats = (size_t)(L - 2) * 4 * 13312;
      if ((rc = upload<float>(&p->wimg, nullptr, wimg_floats))) { tdb200_plan_destroy(p); return rc; }
      cudaMemset(p->wimg, 0, wimg_floats * sizeof(float));
      p->tc_scratch_per_cta = (long long)2 * (L - 1) * 48 * 104;
      if ((rc = upload<float>(&p->tc_scratch, nullptr, (size_t)p->tc_grid * p->tc_scratch_per_cta))) { tdb200_plan_destroy(p); return rc; }
    }
  }
  p->grad_rows = p->grid + tdb::jet_tc_partial_rows() * p->tc_grid + p->simt_rest_grid + (p->tcs_eligible ? 2 * p->n_sms : 0);
  p->loss_rows = p->grid + p->tc_grid + p->simt_rest_grid + (p->tcs_eligible ? p->n_sms : 0);
  for (const auto& e : p->tc_extra) { p->grad_rows += tdb::jet_tc_partial_rows() * e.grid; p->loss_rows += e.grid; }
  if ((rc = upload<float>(&p->arena, nullptr, a.n_params_pad))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload<float>(&p->arena_t, nullptr, a.n_params_pad))) { tdb200_plan_destroy(p); return rc; }
  {
    const size_t img = (size_t)L * tdb::kMaxW * tdb::kWLd;
    if ((rc = upload<float>(&p->img_f, nullptr, img))) { tdb200_plan_destroy(p); return rc; }
    if ((rc = upload<float>(&p->img_b, nullptr, img))) { tdb200_plan_destroy(p); return rc; }
    cudaMemset(p->img_f, 0, img * sizeof(float));
    cudaMemset(p->img_b, 0, img * sizeof(float));
    a.img_f = p->img_f;
    a.img_b = p->img_b;
  }
  if ((rc = upload<float>(&p->part_grad, nullptr, (size_t)p->grad_rows * a.n_params_pad))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload<double>(&p->part_loss, nullptr, (size_t)p->loss_rows * n_slots))) { tdb200_plan_destroy(p); return rc; }
  if ((rc = upload<float>(&p->scratch, nullptr, (size_t)p->grid * a.scratch_per_cta))) { tdb200_plan_destroy(p); return rc; }
  a.arena = p->arena;
  a.arena_t = p->arena_t;
  a.segs = p->d_segs;
  a.seg_tile_begin = p->d_seg_tile_begin;
  a.terms = p->d_terms;
  a.factors = p->d_factors;
  a.n_terms = n_terms;
  a.n_factors = n_factors;
  a.comb = p->d_comb;
  a.slot_scale = p->d_slot_scale;
  a.part_grad = p->part_grad;
  a.part_loss = p->part_loss;
  a.scratch = p->scratch;
  *out = p;
  return TDB200_OK;
}

int tdb200_plan_set_points(tdb200_plan* p, const float* pts_dev, int64_t n_pts, const float* targets_dev,
                           int64_t n_targets, const float* coeffs_dev, int64_t n_coeffs) {
  if (!p || !pts_dev) return fail(TDB200_ERR_INVALID, "null argument");
  int64_t need = 0;
  for (const auto& sg : p->segs) {
    const int64_t end = sg.pts_off + sg.n_groups * sg.K;
    need = end > need ? end : need;
    if (sg.tgt_off >= 0 && (!targets_dev || sg.tgt_off + sg.n_groups * sg.n_cols > n_targets))
      return fail(TDB200_ERR_INVALID, "targets buffer too small");
  }
  if (need > n_pts) return fail(TDB200_ERR_INVALID, "points buffer too small");
  (void)n_coeffs;
  p->pts = pts_dev;
  p->targets = targets_dev;
  p->coeffs = coeffs_dev;
  p->n_pts = n_pts;
  p->args.pts = pts_dev;
  p->args.targets = targets_dev;
  p->args.coeffs = coeffs_dev;
  return TDB200_OK;
}

int tdb200_plan_set_slots(tdb200_plan* p, const double* slot_lambda, const double* slot_len) {
  if (!p || !slot_lambda || !slot_len) return fail(TDB200_ERR_INVALID, "null argument");
  std::vector<float> scale(p->n_slots);
  for (int s = 0; s < p->n_slots; ++s) {
    if (!(slot_len[s] > 0)) return fail(TDB200_ERR_INVALID, "slot_len must be positive");
    scale[s] = (float)(slot_lambda[s] / slot_len[s]);
  }
  CU(cudaSetDevice(p->device));
  // launches of this plan run on caller streams that need not synchronise with the legacy default stream (torch pool
  // streams, the plan's side stream): wait for everything in flight on the device before the tables change, so that no
  // running kernel can read a mix of old and new scales (lambdas change rarely: AdaptiveLambda, every N epochs)
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(p->d_slot_scale, scale.data(), p->n_slots * sizeof(float), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(p->d_slot_lambda, slot_lambda, p->n_slots * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(p->d_slot_len, slot_len, p->n_slots * sizeof(double), cudaMemcpyHostToDevice));
  return TDB200_OK;
}

int tdb200_plan_set_row_weights(tdb200_plan* p, const float* weights_dev) {
  if (!p) return fail(TDB200_ERR_INVALID, "null plan");
  p->args.row_weight = weights_dev;                      // NULL: unweighted (default)
  return TDB200_OK;
}

int tdb200_plan_set_field_seeds(tdb200_plan* p, const float* seeds_dev) {
  if (!p) return fail(TDB200_ERR_INVALID, "null plan");
  p->args.field_seed = seeds_dev;                        // NULL: loss gradient (default)
  return TDB200_OK;
}

int tdb200_comm_create(const void* unique_id, int32_t rank, int32_t world, int32_t device, tdb200_comm** out) {
  if (!out || !unique_id || world < 1 || rank < 0 || rank >= world) return fail(TDB200_ERR_INVALID, "bad communicator arguments");
  CU(cudaSetDevice(device));
  void* c = nullptr;
  const int rc = tdb::comm_create(unique_id, rank, world, &c);
  if (rc != TDB200_OK) return rc;
  *out = reinterpret_cast<tdb200_comm*>(c);
  return TDB200_OK;
}

void tdb200_comm_destroy(tdb200_comm* comm) { tdb::comm_destroy(comm); }

int tdb200_plan_set_peer(tdb200_plan* p, tdb200_peer* peer) {
  if (!p) return fail(TDB200_ERR_INVALID, "null plan");
  p->peer = peer;                                        // borrowed, like the communicator
  return TDB200_OK;
}
int tdb200_plan_set_comm(tdb200_plan* p, tdb200_comm* comm) {
  if (!p) return fail(TDB200_ERR_INVALID, "null plan");
  p->comm = comm;                                        // borrowed: the caller keeps the communicator alive
  return TDB200_OK;
}

int tdb200_plan_set_impl(tdb200_plan* p, int32_t impl) {
  if (!p) return fail(TDB200_ERR_INVALID, "null plan");
  if (impl < 0 || impl > 3) return fail(TDB200_ERR_INVALID, "impl must be 0 (auto), 1 (SIMT), 2 (tcgen05, dW in TMEM) or 3 (tcgen05, streamed dW)");
  if (impl == 3 && !p->tcs_eligible)
    return fail(TDB200_ERR_INVALID, "streamed tcgen05 path needs equal hidden widths <= 104, >= 1 W x W layer and an identity interior segment with pure partials along <= 3 axes");
  if (impl == 2 && !p->tc_eligible)
    return fail(TDB200_ERR_INVALID, "tcgen05 path needs equal hidden widths <= 104, 1 or 2 W x W layers and an identity interior segment with pure partials along <= 3 axes");
  p->impl = impl;
  return TDB200_OK;
}

int64_t tdb200_plan_out_size(const tdb200_plan* p) { return p ? 2 + p->n_slots + p->args.n_params : 0; }
int64_t tdb200_plan_n_params(const tdb200_plan* p) { return p ? p->args.n_params : 0; }
int64_t tdb200_plan_n_fields(const tdb200_plan* p) { return p ? p->n_fields : 0; }
static bool use_tcs(const tdb200_plan* p) {
  if (!p->tcs_eligible || p->impl == 1 || p->impl == 2) return false;
  if (p->impl == 3) return true;
  if (getenv("TDB200_AUTO_TCS")) return atoi(getenv("TDB200_AUTO_TCS")) != 0 && p->segs[0].n_groups >= 4096;
  // >= 4096 rows.  Since the value-row boundary segments ride in the interior launch the pair also wins on short launches
  // (BASELINE config 1, 9 801 + 303 points: 0.119 ms against 0.141 ms for the TMEM-dW kernel); TDB200_AUTO_TCS=0 or impl = 2
  // select jet_tc_kernel.
  return p->segs[0].n_groups >= 4096;
}
static bool use_tc(const tdb200_plan* p) {
  if (!p->tc_eligible || p->impl == 1 || use_tcs(p)) return false;
  return p->impl == 2 || p->segs[0].n_groups >= 4096;
}
// Boundary segments with a launch pair of their own, grouped: consecutive segments of one jet signature whose loss slots
// span < TDB200_MAX_COLS (and whose streams fit the chunk buffers) share ONE jet_tcs + wgrad launch pair.  -> [begin, end)
// index ranges into p->tcs_extra
static std::vector<std::pair<int, int>> tcs_extra_groups(const tdb200_plan* p) {
  std::vector<std::pair<int, int>> out;
  const int Wp = (p->args.widths[1] + 3) / 4 * 4;
  for (size_t i = 0; i < p->tcs_extra.size();) {
    const auto& e0 = p->tcs_extra[i];
    const int Qe = (tdb::jet_tc_columns_per_part(e0.sig[0], e0.sig[1], e0.sig[2]) + 3) / 4;
    const long long per_tile = 4LL * Qe * Wp * 4;
    long long tiles = 0;
    int lo = 1 << 30, hi = -1, n = 0;
    size_t j = i;
    for (; j < p->tcs_extra.size() && n < tdb::kTcsMaxSegs; ++j, ++n) {
      const auto& e = p->tcs_extra[j];
      const tdb200_segment& sg = p->segs[e.seg];
      if (e.sig[0] != e0.sig[0] || e.sig[1] != e0.sig[1] || e.sig[2] != e0.sig[2]) break;
      const tdb200_segment& s0 = p->segs[e0.seg];     // the launch propagates the jets of its FIRST segment's directions
      if (j > i && (sg.n_dirs != s0.n_dirs || memcmp(sg.dir_vec, s0.dir_vec, sizeof(sg.dir_vec)) != 0 ||
                    memcmp(sg.dir_order, s0.dir_order, sizeof(sg.dir_order)) != 0)) break;
      int l2 = lo, h2 = hi;
      for (int c = 0; c < sg.n_cols; ++c) { l2 = std::min(l2, sg.col_slot[c]); h2 = std::max(h2, sg.col_slot[c]); }
      if (j > i && (h2 - l2 >= TDB200_MAX_COLS || getenv("TDB200_TCS_NO_MERGE") ||
                    (p->tcs_stream_stride > 0 && (tiles + e.tiles) * per_tile > p->tcs_stream_stride))) break;
      lo = l2; hi = h2;
      tiles += e.tiles;
    }
    out.push_back({(int)i, (int)j});
    i = j;
  }
  return out;
}
static int tcs_chunks(const tdb200_plan* p) {
  return p->tcs_chunk_tiles > 0 ? (p->tcs_tiles + p->tcs_chunk_tiles - 1) / p->tcs_chunk_tiles : 1;
}
// stream / scratch buffers of the streamed path, sized on first use (never inside a stream capture: warm up first)
static int ensure_tcs_buffers(tdb200_plan* p) {
  if (p->tcs_ys) return TDB200_OK;
  const tdb::JetArgs& a = p->args;
  const int NM = a.n_layers - 2, Wp = (a.widths[1] + 3) / 4 * 4;
  const int P = tdb::jet_tc_points_per_tile(p->tc_sig[0], p->tc_sig[1], p->tc_sig[2]);
  const int J = 1 + p->tc_sig[0] + p->tc_sig[1] + p->tc_sig[2];
  const double budget = (getenv("TDB200_TCS_SCRATCH_MB") ? atof(getenv("TDB200_TCS_SCRATCH_MB")) : 8192.0) * 1048576.0;
  const int Q = (tdb::jet_tc_columns_per_part(p->tc_sig[0], p->tc_sig[1], p->tc_sig[2]) + 3) / 4;
  const double per_tile = 4.0 * Q * Wp * 16.0 * 2.0 * NM;
  long long ct = (long long)(budget / per_tile);
  const int pairs = 2 * p->n_sms;
  ct = ct / pairs * pairs;
  if (ct < pairs) ct = pairs;
  if (ct > p->tcs_tiles) ct = p->tcs_tiles;
  p->tcs_chunk_tiles = (int)ct;
  p->tcs_stream_stride = (long long)ct * 4 * Q * Wp * 4;
  (void)P; (void)J;
  int rc;
  if ((rc = upload<float>(&p->tcs_ys, nullptr, (size_t)p->tcs_stream_stride * NM))) return rc;
  if ((rc = upload<float>(&p->tcs_gs, nullptr, (size_t)p->tcs_stream_stride * NM))) return rc;
  if ((rc = upload<float>(&p->tcs_zsave, nullptr, (size_t)p->n_sms * 2 * NM * 512 * 16))) return rc;
  return TDB200_OK;
}
int32_t tdb200_plan_kernel_path(const tdb200_plan* p) {
  if (!p) return 0;
  return use_tcs(p) ? 3 : use_tc(p) ? 2 : 1;
}
int32_t tdb200_plan_launches_per_call(const tdb200_plan* p) {
  if (!p) return 0;
  if (use_tcs(p)) return 2 + 2 * tcs_chunks(p) + 2 * (int)tcs_extra_groups(p).size() + (p->simt_rest_all_tiles > 0 ? 1 : 0);
  if (!use_tc(p)) return 3;
  return 3 + (int)p->tc_extra.size() + (p->simt_rest_tiles > 0 ? 1 : 0);
}

static int run(tdb200_plan* p, const float* const* params, float* fields, float* out, int do_grad, void* stream) {
  if (!p || !params) return fail(TDB200_ERR_INVALID, "null argument");
  if (!p->pts) return fail(TDB200_ERR_INVALID, "tdb200_plan_set_points was not called");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(p->device));
  tdb::PackArgs pk{};
  const tdb::JetArgs& a = p->args;
  pk.n_layers = a.n_layers;
  for (int l = 0; l <= a.n_layers; ++l) pk.widths[l] = a.widths[l];
  for (int l = 0; l < a.n_layers; ++l) {
    pk.w_off[l] = a.w_off[l];
    pk.b_off[l] = a.b_off[l];
    pk.W[l] = params[2 * l];
    pk.b[l] = params[2 * l + 1];
    if (!pk.W[l] || !pk.b[l]) return fail(TDB200_ERR_INVALID, "null parameter pointer");
  }
  pk.n_net_params = a.n_net_params;
  pk.n_cparams = a.n_cparams;
  for (int i = 0; i < a.n_cparams; ++i) pk.c[i] = params[2 * a.n_layers + i];
  pk.arena = p->arena;
  pk.arena_t = p->arena_t;
  pk.img_f = p->img_f;
  pk.img_b = p->img_b;
  pk.wimg = (use_tcs(p) || use_tc(p)) ? p->wimg : nullptr;     // tensor-core weight images in the same launch
  CU(tdb::launch_pack_params(pk, s));
  tdb::JetArgs call = a;
  call.fields = fields;
  call.do_grad = do_grad;
  int grad_rows = p->grid, loss_rows = p->grid;
  if (use_tcs(p)) {
    { const int rc = ensure_tcs_buffers(p); if (rc != TDB200_OK) return rc; }
    const int NM = a.n_layers - 2, Wp = (a.widths[1] + 3) / 4 * 4;
    const int Q = (tdb::jet_tc_columns_per_part(p->tc_sig[0], p->tc_sig[1], p->tc_sig[2]) + 3) / 4;
    // boundary rows: SIMT kernel on a side stream, on the SMs the persistent interior grid leaves free
    const double tcs_us = 30.0 + 9.0 * ((p->tcs_tiles + p->tc_grid - 1) / p->tc_grid) * (1.0 + 0.5 * NM);
    int rest_ctas = p->simt_rest_all_tiles < p->n_sms ? p->simt_rest_all_tiles : p->n_sms, reserve = 0;
    bool fork = false;
    if (p->simt_rest_all_tiles > 0 && p->tc_grid == p->n_sms && !getenv("TDB200_NO_OVERLAP")) {
      int k = (int)ceil(p->simt_rest_all_tiles * 170.0 / (tcs_us > 50.0 ? tcs_us : 50.0));
      k = k < 1 ? 1 : k;
      if (k <= p->n_sms / 8) { rest_ctas = k; reserve = k; fork = true; }
    }
    const int gA = p->tc_grid - reserve;
    const int gG = gA;
    const int rowsA = tdb::jet_tc_partial_rows() * gA;
    grad_rows = rowsA;                     // the weight-gradient GEMM writes its dW blocks into the SAME partial rows
    loss_rows = gA;
    cudaStream_t ss = s;
    if (fork) {
      if (!p->side) {
        CU(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
      }
      CU(cudaEventRecord(p->ev_fork, s));
      CU(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
      ss = p->side;
    }
    auto launch_rest = [&]() -> int {
      if (p->simt_rest_all_tiles > 0) {
        tdb::JetArgs rest = call;
        rest.seg_tile_begin = p->d_seg_tile_begin_rest_all;
        rest.n_tiles = p->simt_rest_all_tiles;
        rest.part_grad = p->part_grad + (size_t)grad_rows * a.n_params_pad;
        rest.part_loss = p->part_loss + (size_t)loss_rows * p->n_slots;
        CU(tdb::launch_jet_simt(rest, rest_ctas, ss));
      }
      return TDB200_OK;
    };
    if (fork) {
      const int rc = launch_rest();
      if (rc != TDB200_OK) return rc;
      CU(cudaEventRecord(p->ev_join, p->side));
    }
    tdb::JetArgs tc = call;
    tc.seg_tile_begin = p->d_seg_tile_begin_tc;
    long long* dbg = nullptr;
    const size_t dbg_n = 16 * (size_t)p->tc_grid + 512;      // + event trace of one iteration of CTA 0 (TDB_TC_TIMING builds)
    if (getenv("TDB200_TC_TIMING")) { cudaMalloc(&dbg, sizeof(long long) * dbg_n); cudaMemset(dbg, 0, sizeof(long long) * dbg_n); }
    tc.dbg = dbg;
    tdb::TcsArgs xa{};
    xa.wimg = p->wimg; xa.ys = p->tcs_ys; xa.gs = p->tcs_gs; xa.stream_stride = p->tcs_stream_stride;
    xa.zsave = p->tcs_zsave; xa.Wp = Wp;
    xa.n_msegs = (int)p->tcs_mseg.size();
    for (int m = 0; m < xa.n_msegs; ++m) { xa.mseg_index[m] = p->tcs_mseg[m]; xa.mseg_tile_begin[m] = p->tcs_mseg_tile_begin[m]; }
    xa.mseg_tile_begin[xa.n_msegs] = p->tcs_tiles;
    xa.term_end = p->tcs_term_end;
    xa.slot_base = p->tcs_slot_base;
    tdb::WgradArgs wa{};
    wa.gs = p->tcs_gs; wa.ys = p->tcs_ys; wa.stream_stride = p->tcs_stream_stride;
    wa.W = a.widths[1]; wa.Wp = Wp; wa.n_mma = NM; wa.kb = tdb::wgrad_kb(); wa.splits = gG / NM > 0 ? gG / NM : 1;
    wa.part = p->part_grad; wa.n_params_pad = a.n_params_pad;         // CTA b -> row b (zero-filled by jet_tcs CTA b)
    for (int t = 1; t <= NM; ++t) wa.w_off[t - 1] = a.w_off[t];
    if (gG < NM) return fail(TDB200_ERR_INVALID, "streamed tcgen05 path: fewer CTAs than W x W layers");
    int chunk = 0;
    for (int t0 = 0; t0 < p->tcs_tiles; t0 += p->tcs_chunk_tiles, ++chunk) {
      const int t1 = t0 + p->tcs_chunk_tiles < p->tcs_tiles ? t0 + p->tcs_chunk_tiles : p->tcs_tiles;
      xa.tile0 = t0; xa.tile1 = t1; xa.zero_partials = chunk == 0;
      CU(tdb::launch_jet_tcs(tc, xa, p->tc_sig[0], p->tc_sig[1], p->tc_sig[2], gA, s));
      if (do_grad) {
        wa.total4 = (long long)(t1 - t0) * 4 * Q * Wp;
        wa.accumulate = chunk > 0;
        CU(tdb::launch_wgrad_gemm(wa, gG, s));
      }
      if (dbg && chunk == 0) {
        std::vector<long long> h(16 * p->tc_grid);
        cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const int tiles_per_cta = (t1 - t0 + gA - 1) / gA;
        fprintf(stderr, "[tdb200 tcs timing] cycles per tile (CTA 0, %d tiles; 0-11 epilogue thread 0, 12-15 MMA warp):", tiles_per_cta);
        for (int i = 0; i < 16; ++i) fprintf(stderr, " p%d=%lld", i, h[i] / tiles_per_cta);
        fprintf(stderr, "\n");
        if (getenv("TDB200_TC_TRACE")) {                 // merged event list of iteration 40 of CTA 0: (cycle, role, mark)
          std::vector<long long> tr(512);
          cudaMemcpy(tr.data(), dbg + 16 * p->tc_grid, 512 * sizeof(long long), cudaMemcpyDeviceToHost);
          std::vector<std::pair<long long, int>> ev;
          for (int r = 0; r < 2; ++r)
            for (int i = 0; i < 256; ++i)
              if (tr[r * 256 + i]) ev.push_back({tr[r * 256 + i] >> 8, r * 1000 + (int)(tr[r * 256 + i] & 255)});
          std::sort(ev.begin(), ev.end());
          for (const auto& e : ev)
            fprintf(stderr, "[tdb200 tcs trace] %8lld %s mark %d\n", e.first - ev[0].first, e.second >= 1000 ? "            mma" : "epi", e.second % 1000);
        }
      }
    }
    if (dbg) cudaFree(dbg);
    // boundary segments of identity rows: the same two kernels in stream order (the stream buffers are reused), adding
    // to the partial rows of the interior launches
    // (consecutive segments of one jet signature - e.g. the eight Dirichlet conditions of the Navier-Stokes config - share
    // ONE launch pair: jet_tcs_kernel walks several segments, see TcsArgs::mseg_*)
    for (const auto& grp : tcs_extra_groups(p)) {
      const auto& e0 = p->tcs_extra[grp.first];
      const int Qe = (tdb::jet_tc_columns_per_part(e0.sig[0], e0.sig[1], e0.sig[2]) + 3) / 4;
      const long long per_tile = 4LL * Qe * Wp * 4;
      tdb::JetArgs xe = call;
      xe.row_weight = nullptr;
      xe.dbg = nullptr;
      xa.n_msegs = 0; xa.term_end = 0;
      int tiles = 0, slot_lo = 1 << 30;
      for (int j = grp.first; j < grp.second; ++j) {
        const auto& e = p->tcs_extra[j];
        const tdb200_segment& sg = p->segs[e.seg];
        for (int c = 0; c < sg.n_cols; ++c) slot_lo = std::min(slot_lo, sg.col_slot[c]);
        xa.mseg_index[xa.n_msegs] = e.seg;
        xa.mseg_tile_begin[xa.n_msegs] = tiles;
        ++xa.n_msegs;
        tiles += e.tiles;
        xa.term_end = std::max(xa.term_end, (int)sg.col_term_end[sg.n_cols - 1]);
      }
      xa.mseg_tile_begin[xa.n_msegs] = tiles;
      xa.slot_base = slot_lo;
      if ((long long)tiles * per_tile > p->tcs_stream_stride) return fail(TDB200_ERR_INVALID, "streamed tcgen05 path: a boundary segment exceeds the stream chunk");
      xa.tile0 = 0; xa.tile1 = tiles; xa.zero_partials = 0;
      const int ge = tiles < gA ? tiles : gA;
      CU(tdb::launch_jet_tcs(xe, xa, e0.sig[0], e0.sig[1], e0.sig[2], ge, s));
      if (do_grad) {
        wa.total4 = (long long)tiles * 4 * Qe * Wp;
        wa.kb = tdb::wgrad_kb();
        wa.accumulate = 1;
        CU(tdb::launch_wgrad_gemm(wa, gG, s));
      }
    }
    if (fork) {
      CU(cudaStreamWaitEvent(s, p->ev_join, 0));
    } else {
      const int rc = launch_rest();
      if (rc != TDB200_OK) return rc;
    }
    if (p->simt_rest_all_tiles > 0) { grad_rows += rest_ctas; loss_rows += rest_ctas; }
  } else if (use_tc(p)) {
    tdb::JetArgs tc = call;
    tc.seg_tile_begin = p->d_seg_tile_begin_tc;
    tc.n_tiles = p->tc_tiles;
    tc.scratch = p->tc_scratch;
    tc.scratch_per_cta = p->tc_scratch_per_cta;
    long long* dbg = nullptr;
    if (getenv("TDB200_TC_TIMING")) { cudaMalloc(&dbg, sizeof(long long) * 16 * p->tc_grid); }
    tc.dbg = dbg;
    // The boundary rows run NEXT TO the interior tcgen05 launch, on a side stream and on SMs the persistent interior
    // grid leaves free: small tcgen05 launches for the identity segments, then the SIMT kernel for the rest (a few tiles
    // of ~170 us each) when few CTAs finish it within the interior launch's own time.
    const double tc_us = 25.0 + 15.0 * ((p->tc_tiles + p->tc_grid - 1) / p->tc_grid);
    const bool may_fork = p->tc_grid == p->n_sms && !dbg && !getenv("TDB200_NO_OVERLAP");
    int rest_ctas = p->simt_rest_grid, reserve = 0, extra_cap = 16;
    bool fork = false;
    if (may_fork) {
      // fewest side CTAs that still finish inside the interior launch's own time (every reserved SM is taken from it)
      auto extras_us = [&](int g) {
        double us = 0.0;
        for (const auto& e : p->tc_extra) { const int ge = e.grid < g ? e.grid : g; us += 30.0 + 15.0 * ((e.tiles + ge - 1) / ge); }
        return us;
      };
      int g = 1;
      while (g < 4 && extras_us(g) > 0.9 * tc_us) ++g;
      const double side_us = extras_us(g);
      if (!p->tc_extra.empty()) { extra_cap = g; reserve = g; }
      if (p->simt_rest_tiles > 0) {
        const double left = tc_us - side_us > 50.0 ? tc_us - side_us : 50.0;
        int k = (int)ceil(p->simt_rest_tiles * 170.0 / left);
        k = k < 1 ? 1 : k;
        if (k <= p->n_sms / 8) { rest_ctas = k; reserve = k > reserve ? k : reserve; fork = true; }
        else fork = false;                                      // too much SIMT work to hide: everything in stream order
      } else {
        fork = !p->tc_extra.empty();
      }
      if (!fork) { reserve = 0; extra_cap = 16; }
    }
    const int tc_grid = p->tc_grid - reserve;
    grad_rows = tdb::jet_tc_partial_rows() * tc_grid;
    loss_rows = tc_grid;
    cudaStream_t ss = s;
    if (fork) {
      if (!p->side) {
        CU(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
      }
      CU(cudaEventRecord(p->ev_fork, s));
      CU(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
      ss = p->side;
    }
    int side_grad_rows = 0, side_loss_rows = 0;                  // rows of the side launches follow the interior rows
    auto launch_side = [&]() -> int {
      for (const auto& e : p->tc_extra) {
        tdb::JetArgs x = call;
        x.segs = a.segs + e.seg;
        x.n_tiles = e.tiles;
        x.row_weight = nullptr;
        x.dbg = nullptr;
        x.part_grad = p->part_grad + (size_t)(grad_rows + side_grad_rows) * a.n_params_pad;
        x.part_loss = p->part_loss + (size_t)(loss_rows + side_loss_rows) * p->n_slots;
        const int ge = e.grid < extra_cap ? e.grid : extra_cap;
        CU(tdb::launch_jet_tc(x, p->wimg, e.sig[0], e.sig[1], e.sig[2], ge, ss));
        side_grad_rows += tdb::jet_tc_partial_rows() * ge;
        side_loss_rows += ge;
      }
      if (p->simt_rest_tiles > 0) {
        tdb::JetArgs rest = call;
        rest.seg_tile_begin = p->d_seg_tile_begin_rest;
        rest.n_tiles = p->simt_rest_tiles;
        rest.part_grad = p->part_grad + (size_t)(grad_rows + side_grad_rows) * a.n_params_pad;
        rest.part_loss = p->part_loss + (size_t)(loss_rows + side_loss_rows) * p->n_slots;
        CU(tdb::launch_jet_simt(rest, rest_ctas, ss));
        side_grad_rows += rest_ctas;
        side_loss_rows += rest_ctas;
      }
      return TDB200_OK;
    };
    if (fork) {                                                 // enqueued first: the side CTAs take their SMs first
      const int rc = launch_side();
      if (rc != TDB200_OK) return rc;
      CU(cudaEventRecord(p->ev_join, p->side));
    }
    CU(tdb::launch_jet_tc(tc, p->wimg, p->tc_sig[0], p->tc_sig[1], p->tc_sig[2], tc_grid, s));
    if (dbg) {
      std::vector<long long> h(16 * p->tc_grid);
      cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      cudaFree(dbg);
      const int tiles_per_cta = (p->tc_tiles + p->tc_grid - 1) / p->tc_grid;
      fprintf(stderr, "[tdb200 tc timing] cycles per tile (thread 0 of CTA 0, %d tiles):", tiles_per_cta);
      for (int i = 0; i < 16; ++i) fprintf(stderr, " p%d=%lld", i, h[i] / tiles_per_cta);
      fprintf(stderr, "\n");
    }
    if (fork) {
      CU(cudaStreamWaitEvent(s, p->ev_join, 0));
    } else {
      const int rc = launch_side();
      if (rc != TDB200_OK) return rc;
    }
    grad_rows += side_grad_rows;
    loss_rows += side_loss_rows;
  } else {
    CU(tdb::launch_jet_simt(call, p->grid, s));
  }
  if (out) {
    CU(tdb::launch_reduce_partials(p->part_grad, do_grad ? grad_rows : 0, p->part_loss, loss_rows,
                                   do_grad ? a.n_params : 0, a.n_params_pad, p->n_slots, p->d_slot_lambda,
                                   p->d_slot_len, out, s));
    if (p->peer) {      // one box: sum over peer memory (csrc/peer.cu), one small kernel of the library
      const int rc = tdb200_peer_allreduce_vec(p->peer, out, (int64_t)(2 + p->n_slots + (do_grad ? a.n_params : 0)), s);
      if (rc != TDB200_OK) return rc;
    } else
    if (p->comm) {      // ranks hold row blocks with global denominators: the per-rank vectors simply add
      const int rc = tdb::comm_all_reduce_sum(p->comm, out, (size_t)(2 + p->n_slots + (do_grad ? a.n_params : 0)), s);
      if (rc != TDB200_OK) return rc;
    }
  }
  return TDB200_OK;
}

int tdb200_loss_grad(tdb200_plan* p, const float* const* params_dev, float* out_dev, void* stream) {
  if (!out_dev) return fail(TDB200_ERR_INVALID, "null output");
  return run(p, params_dev, nullptr, out_dev, 1, stream);
}

int tdb200_eval_fields(tdb200_plan* p, const float* const* params_dev, float* fields_dev, float* out_dev,
                       void* stream) {
  if (!fields_dev) return fail(TDB200_ERR_INVALID, "null fields buffer");
  return run(p, params_dev, fields_dev, out_dev, 0, stream);
}

int64_t tdb200_plan_n_params_pad(const tdb200_plan* p) { return p ? p->args.n_params_pad : 0; }

int tdb200_jacobian_rows(tdb200_plan* p, const float* const* params, int32_t segment, int32_t col, float* rows_dev,
                         void* stream) {
  if (!p || !params || !rows_dev) return fail(TDB200_ERR_INVALID, "null argument");
  if (!p->pts) return fail(TDB200_ERR_INVALID, "tdb200_plan_set_points was not called");
  if (segment < 0 || segment >= (int)p->segs.size()) return fail(TDB200_ERR_INVALID, "segment out of range");
  if (col < 0 || col >= p->segs[segment].n_cols) return fail(TDB200_ERR_INVALID, "residual column out of range");
  if (p->segs[segment].n_groups > 0x7fffffffLL) return fail(TDB200_ERR_INVALID, "too many rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(p->device));
  tdb::PackArgs pk{};
  const tdb::JetArgs& a = p->args;
  pk.n_layers = a.n_layers;
  for (int l = 0; l <= a.n_layers; ++l) pk.widths[l] = a.widths[l];
  for (int l = 0; l < a.n_layers; ++l) {
    pk.w_off[l] = a.w_off[l];
    pk.b_off[l] = a.b_off[l];
    pk.W[l] = params[2 * l];
    pk.b[l] = params[2 * l + 1];
    if (!pk.W[l] || !pk.b[l]) return fail(TDB200_ERR_INVALID, "null parameter pointer");
  }
  pk.n_net_params = a.n_net_params;
  pk.n_cparams = a.n_cparams;
  for (int i = 0; i < a.n_cparams; ++i) pk.c[i] = params[2 * a.n_layers + i];
  pk.arena = p->arena; pk.arena_t = p->arena_t; pk.img_f = p->img_f; pk.img_b = p->img_b;
  CU(tdb::launch_pack_params(pk, s));
  tdb::JetArgs call = a;
  call.fields = nullptr; call.row_weight = nullptr; call.field_seed = nullptr; call.dbg = nullptr;
  call.do_grad = 1;
  call.jac_rows = rows_dev; call.jac_seg = segment; call.jac_col = col;
  call.n_tiles = (int)p->segs[segment].n_groups;
  if (call.n_tiles == 0) return TDB200_OK;
  CU(tdb::launch_jet_simt(call, call.n_tiles < p->grid ? call.n_tiles : p->grid, s));
  return TDB200_OK;
}

void tdb200_plan_destroy(tdb200_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaFree(p->d_segs); cudaFree(p->d_seg_tile_begin); cudaFree(p->d_terms); cudaFree(p->d_factors);
  cudaFree(p->d_comb); cudaFree(p->d_slot_scale); cudaFree(p->d_slot_lambda); cudaFree(p->d_slot_len);
  cudaFree(p->arena); cudaFree(p->arena_t); cudaFree(p->img_f); cudaFree(p->img_b);
  cudaFree(p->d_seg_tile_begin_tc); cudaFree(p->d_seg_tile_begin_rest); cudaFree(p->d_seg_tile_begin_rest_all); cudaFree(p->tcs_ys); cudaFree(p->tcs_gs); cudaFree(p->tcs_zsave); cudaFree(p->wimg); cudaFree(p->tc_scratch); cudaFree(p->part_grad); cudaFree(p->part_loss); cudaFree(p->scratch);
  if (p->side) { cudaStreamDestroy(p->side); cudaEventDestroy(p->ev_fork); cudaEventDestroy(p->ev_join); }
  delete p;
}

}  // extern "C"
