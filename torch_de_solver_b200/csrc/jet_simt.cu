// Fused Taylor-jet MLP loss + gradient kernel, SIMT fp32 version (sm_100a).
//
// One persistent CTA per SM walks over tiles of collocation points.  For every tile it runs, entirely in
// shared memory / L2-resident scratch:
//   forward : layer 0 (K = d) -> tanh jets -> [GEMM -> tanh jets]* -> last layer -> operator terms
//             -> residual per row -> loss partial
//   backward: residual adjoint -> last layer -> [tanh-jet adjoint -> dW, db -> data GEMM]* -> layer 0
// and accumulates the parameter gradient into a per-CTA partial that a second kernel reduces in a fixed
// order (deterministic, no float atomics on the gradient).
//
// Replaces, per optimiser step: the 2^k shifted MLP forwards of Derivative_NN (tedeous/derivative.py:47-51),
// the nested autograd.grad calls of Derivative_autograd (derivative.py:92-97), Operator.apply_operator
// (eval.py:143-165), Bounds.apply_bcs (eval.py:433-461), Losses._default_loss (losses.py:84-135) and
// loss.backward() (optimizers/closure.py:60).
//
// Activation layout: act[k][r], k = neuron, r = c * P + p (jet channel c, point p), row stride kLd.
#include "common.cuh"

namespace tdb {

// ------------------------------------------------------------------------------------------------
// parameter packing: gather the torch parameter tensors into one arena (gradient layout) and write
// the transposed weights the forward GEMMs stream.
// ------------------------------------------------------------------------------------------------
__global__ void pack_params_kernel(PackArgs a) {
  pdl_launch_dependents();             // the fused kernel behind may run its set-up next to this launch
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  for (int l = 0; l < a.n_layers; ++l) {
    const int in = a.widths[l], out = a.widths[l + 1];
    const float* __restrict__ W = a.W[l];
    for (int i = tid; i < in * out; i += nth) {
      const float w = W[i];
      a.arena[a.w_off[l] + i] = w;
      const int n = i / in, k = i - n * in;
      a.arena_t[a.w_off[l] + k * out + n] = w;
      // padded images the GEMM tiles copy verbatim: [reduction index][8 warp groups x 16]
      const int tf = (out + 7) / 8 <= 4 ? 4 : (out + 7) / 8 <= 8 ? 8 : (out + 7) / 8 <= 13 ? 13 : 16;
      const int tb = (in + 7) / 8 <= 4 ? 4 : (in + 7) / 8 <= 8 ? 8 : (in + 7) / 8 <= 13 ? 13 : 16;
      a.img_f[(size_t)l * kMaxW * kWLd + k * kWLd + (n / tf) * 16 + n % tf] = w;
      a.img_b[(size_t)l * kMaxW * kWLd + n * kWLd + (k / tb) * 16 + k % tb] = w;
      if (a.wimg && l >= 1 && l <= a.n_layers - 2) {        // tensor-core images: W hi | W lo | W^T hi | W^T lo
        float* hi = a.wimg + (size_t)(l - 1) * 4 * kTcWFloats;
        const float h = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xFFFFE000u);      // tf32, round to nearest
        const int o = sw_off(n, k, kTcWRows), ot = sw_off(k, n, kTcWRows);
        hi[o] = h;
        hi[kTcWFloats + o] = w - h;
        hi[2 * kTcWFloats + ot] = h;
        hi[3 * kTcWFloats + ot] = w - h;
      }
    }
    for (int i = tid; i < out; i += nth) a.arena[a.b_off[l] + i] = a.b[l][i];
  }
  for (int i = tid; i < a.n_cparams; i += nth) a.arena[a.n_net_params + i] = a.c[i][0];
}

cudaError_t launch_pack_params(const PackArgs& a, cudaStream_t s) {
  pack_params_kernel<<<148, 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------------
struct Smem {
  float* actA;     // [wmax][kLd]
  float* actB;     // [wmax][kLd]
  float* wS;       // [wmax][kWLd]
  float* xS;       // [kRows][4]
  float* uS;       // [n_out][R]    jets of the network outputs
  float* guS;      // [n_out][R]
  float* vS;       // [n_out][M][G] virtual channels (non-identity segments)
  float* gvS;      // [n_out][M][G]
  float* rS;       // [cols][G]     residual adjoint seeds
  float* wlS;      // [n_out][kMaxW] last-layer weights
  float* cgS;      // [kMaxCParams] coefficient-parameter gradient accumulators
  double* lossS;   // [<=32 slots]
};

__host__ __device__ inline size_t smem_floats(int wmax) {
  return (size_t)2 * wmax * kLd + (size_t)wmax * kWLd + kRows * 4 + 4 * kMaxOut * kRows +
         TDB200_MAX_COLS * kRows + kMaxOut * kMaxW + kMaxCParams;
}
size_t jet_simt_smem_bytes(int wmax) { return smem_floats(wmax) * sizeof(float) + 32 * sizeof(double) + 16; }

__device__ inline Smem carve(float* base, int wmax) {
  Smem s;
  s.actA = base;
  s.actB = s.actA + (size_t)wmax * kLd;
  s.wS = s.actB + (size_t)wmax * kLd;
  s.xS = s.wS + (size_t)wmax * kWLd;
  s.uS = s.xS + kRows * 4;
  s.guS = s.uS + kMaxOut * kRows;
  s.vS = s.guS + kMaxOut * kRows;
  s.gvS = s.vS + kMaxOut * kRows;
  s.rS = s.gvS + kMaxOut * kRows;
  s.wlS = s.rS + TDB200_MAX_COLS * kRows;
  s.cgS = s.wlS + kMaxOut * kMaxW;
  float* end = s.cgS + kMaxCParams;
  s.lossS = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(end) + 15) & ~uintptr_t(15));
  return s;
}

// ------------------------------------------------------------------------------------------------
// GEMM building blocks
// ------------------------------------------------------------------------------------------------
// weight tile loader: src is [I][Jd] (reduction index major); column j of warp-group j / TN lands at
// wS[i][ (j / TN) * 16 + j % TN ] so every warp reads its TN columns with 16-byte broadcast loads.
__device__ __forceinline__ void load_weight_image(float* __restrict__ wS, const float* __restrict__ img, int I) {
  const float4* s4 = reinterpret_cast<const float4*>(img);
  float4* d4 = reinterpret_cast<float4*>(wS);
  for (int idx = threadIdx.x; idx < I * (kWLd / 4); idx += kThreads) d4[idx] = __ldg(s4 + idx);
}

// out[j][r] = sum_i in[i][r] * w[i][j]   for r < R (multiple of 4), j < Jd <= 8 * TN
template <int TN>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ in, const float* __restrict__ wS,
                                          float* __restrict__ out, int I, int Jd, int R) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = lane * 4;
  if (r0 >= R || warp * TN >= Jd) return;
  float acc[TN][4];
#pragma unroll
  for (int j = 0; j < TN; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const float* ip = in + r0;
  const float* wp = wS + warp * 16;
  constexpr int NV = (TN + 3) / 4;
#pragma unroll 2
  for (int i = 0; i < I; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(ip + (size_t)i * kLd);
    float w[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 t = *reinterpret_cast<const float4*>(wp + (size_t)i * kWLd + v * 4);
      w[v * 4 + 0] = t.x; w[v * 4 + 1] = t.y; w[v * 4 + 2] = t.z; w[v * 4 + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      acc[j][0] = fmaf(a.x, w[j], acc[j][0]);
      acc[j][1] = fmaf(a.y, w[j], acc[j][1]);
      acc[j][2] = fmaf(a.z, w[j], acc[j][2]);
      acc[j][3] = fmaf(a.w, w[j], acc[j][3]);
    }
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int col = warp * TN + j;
    if (col < Jd)
      *reinterpret_cast<float4*>(out + (size_t)col * kLd + r0) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
  }
}

__device__ __forceinline__ int pick_tn(int jd) { return (jd + 7) / 8; }

__device__ __forceinline__ void gemm_rows_any(const float* in, const float* wS, float* out, int I, int Jd, int R) {
  const int t = pick_tn(Jd);
  if (t <= 4) gemm_rows<4>(in, wS, out, I, Jd, R);
  else if (t <= 8) gemm_rows<8>(in, wS, out, I, Jd, R);
  else if (t <= 13) gemm_rows<13>(in, wS, out, I, Jd, R);
  else gemm_rows<16>(in, wS, out, I, Jd, R);
}

// dst[n * Kd + k] += sum_{r < R} g[n][r] * y[k][r]    (weight gradient; reduction over the tile rows)
template <int TW>
__device__ __forceinline__ void gemm_wgrad(const float* __restrict__ g, const float* __restrict__ y,
                                           float* __restrict__ dst, int Nd, int Kd, int R) {
  const int tn = threadIdx.x >> 4, tk = threadIdx.x & 15;
  if (tn >= Nd && tk >= Kd) return;
  float acc[TW][TW];
#pragma unroll
  for (int a = 0; a < TW; ++a)
#pragma unroll
    for (int b = 0; b < TW; ++b) acc[a][b] = 0.f;
  for (int r = 0; r < R; r += 4) {
    float4 gv[TW], yv[TW];
#pragma unroll
    for (int a = 0; a < TW; ++a) {
      const int n = tn + 16 * a;
      gv[a] = n < Nd ? *reinterpret_cast<const float4*>(g + (size_t)n * kLd + r) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int b = 0; b < TW; ++b) {
      const int k = tk + 16 * b;
      yv[b] = k < Kd ? *reinterpret_cast<const float4*>(y + (size_t)k * kLd + r) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int a = 0; a < TW; ++a)
#pragma unroll
      for (int b = 0; b < TW; ++b) {
        acc[a][b] = fmaf(gv[a].x, yv[b].x, acc[a][b]);
        acc[a][b] = fmaf(gv[a].y, yv[b].y, acc[a][b]);
        acc[a][b] = fmaf(gv[a].z, yv[b].z, acc[a][b]);
        acc[a][b] = fmaf(gv[a].w, yv[b].w, acc[a][b]);
      }
  }
#pragma unroll
  for (int a = 0; a < TW; ++a) {
    const int n = tn + 16 * a;
    if (n >= Nd) continue;
#pragma unroll
    for (int b = 0; b < TW; ++b) {
      const int k = tk + 16 * b;
      if (k < Kd) atomicAdd(dst + (size_t)n * Kd + k, acc[a][b]);   // RED: single owner per address, no stall
    }
  }
}
__device__ __forceinline__ void gemm_wgrad_any(const float* g, const float* y, float* dst, int Nd, int Kd, int R) {
  const int t = (max(Nd, Kd) + 15) / 16;
  if (t <= 2) gemm_wgrad<2>(g, y, dst, Nd, Kd, R);
  else if (t <= 4) gemm_wgrad<4>(g, y, dst, Nd, Kd, R);
  else if (t <= 7) gemm_wgrad<7>(g, y, dst, Nd, Kd, R);
  else gemm_wgrad<8>(g, y, dst, Nd, Kd, R);
}

// first-order input of jet direction v after the first Linear layer: W0[n, :] . v (a column of W0 for a pure partial)
__device__ __forceinline__ float dir_dot(const float* __restrict__ w0_row, const float* v, int d) {
  float s = 0.f;
  for (int ax = 0; ax < d; ++ax) s = fmaf(__ldg(w0_row + ax), v[ax], s);
  return s;
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) jet_simt_kernel(const JetArgs a) {
  extern __shared__ __align__(16) float smem_raw[];
  const Smem sm = carve(smem_raw, a.wmax);
  const int tid = threadIdx.x;
  const int L = a.n_layers;
  const int n_out = a.widths[L];
  // Jacobian-rows mode (a.jac_rows != nullptr, tdb200_jacobian_rows): every tile holds ONE group of segment a.jac_seg and
  // accumulates into its own output row, seeded with 1 on residual column a.jac_col: row r = d field[r, col] / d theta
  // (the per-residual Jacobian the reference's NGD builds with one autograd.grad per point, optimizers/ngd.py:57-77).
  const bool jac = a.jac_rows != nullptr;
  float* my_grad = a.part_grad + (size_t)blockIdx.x * a.n_params_pad;
  float* const my_scratch = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
  const size_t save_stride = (size_t)a.wmax * kRows;       // one saved [w][128] block

  // zero this CTA's partials
  if (a.do_grad && !jac)
    for (int i = tid; i < a.n_params_pad; i += kThreads) my_grad[i] = 0.f;
  if (tid < 32) sm.lossS[tid] = 0.0;
  if (tid < kMaxCParams) sm.cgS[tid] = 0.f;
  // last-layer weights stay resident
  for (int i = tid; i < n_out * a.widths[L - 1]; i += kThreads) {
    const int v = i / a.widths[L - 1], k = i - v * a.widths[L - 1];
    sm.wlS[v * kMaxW + k] = a.arena[a.w_off[L - 1] + i];
  }
  for (int i = tid; i < 2 * a.wmax * kLd; i += kThreads) sm.actA[i] = 0.f;
  __syncthreads();

  int seg_i = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    if (jac) seg_i = a.jac_seg;
    else while (tile >= a.seg_tile_begin[seg_i + 1]) ++seg_i;
    const tdb200_segment& sg = a.segs[seg_i];
    const int K = sg.K, M = sg.M, ncols = sg.n_cols, ndirs = sg.n_dirs;
    int J = 1;
    for (int i = 0; i < ndirs; ++i) J += sg.dir_order[i];
    // points per tile: largest multiple of lcm(4, K) with J * P <= 128 (plan.points_per_tile)
    int step = K;
    if (step % 4) step = (step % 2) ? step * 4 : step * 2;
    const int P = ((kRows / J) / step) * step;
    const int G = P / K;
    const int R = J * P;
    const long long g_first = jac ? (long long)tile : (long long)(tile - a.seg_tile_begin[seg_i]) * G;
    const int g_valid = jac ? 1 : (int)min((long long)G, sg.n_groups - g_first);
    if (jac) {
      my_grad = a.jac_rows + (size_t)tile * a.n_params_pad;
      for (int i = tid; i < a.n_params_pad; i += kThreads) my_grad[i] = 0.f;      // visible after the barrier below
    }
    const int p_valid = g_valid * K;
    const int d = a.d;

    // ---- point coordinates -----------------------------------------------------------------------
    for (int i = tid; i < P * d; i += kThreads) {
      const int p = i / d, ax = i - p * d;
      sm.xS[p * 4 + ax] = p < p_valid ? __ldg(a.pts + (size_t)(sg.pts_off + g_first * K + p) * d + ax) : 0.f;
    }
    __syncthreads();

    // ---- layer 0: K = d, no GEMM; derivative channels start as columns of W0 --------------------
    float* cur = sm.actA;
    float* oth = sm.actB;
    {
      const int w1 = a.widths[1];
      const float* W0 = a.arena + a.w_off[0];
      const float* b0 = a.arena + a.b_off[0];
      float* ysave = my_scratch;                           // layer-0 block: Y only
      for (int idx = tid; idx < w1 * P; idx += kThreads) {
        const int n = idx / P, p = idx - n * P;
        float z0 = __ldg(b0 + n);
        for (int ax = 0; ax < d; ++ax) z0 = fmaf(__ldg(W0 + n * d + ax), sm.xS[p * 4 + ax], z0);
        const float av = tanhf(z0);
        float* row = cur + (size_t)n * kLd;
        row[p] = av;
        if (a.do_grad) ysave[(size_t)n * kRows + p] = av;
        if (J > 1) {
          const TanhF f(av);
          int c = 1;
          for (int i = 0; i < ndirs; ++i) {
            const int o = sg.dir_order[i];
            float z[4] = {dir_dot(W0 + n * d, sg.dir_vec[i], d), 0.f, 0.f, 0.f}, y[4];
            tanh_jet_fwd(f, z, o, y);
            for (int k = 0; k < o; ++k) {
              row[(c + k) * P + p] = y[k];
              if (a.do_grad) ysave[(size_t)n * kRows + (c + k) * P + p] = y[k];
            }
            c += o;
          }
        }
      }
    }
    __syncthreads();

    // ---- hidden layers 1 .. L-2: GEMM + tanh-jet epilogue ---------------------------------------
    for (int l = 1; l <= L - 2; ++l) {
      const int Kd = a.widths[l], Nd = a.widths[l + 1];
      load_weight_image(sm.wS, a.img_f + (size_t)l * kMaxW * kWLd, Kd);
      __syncthreads();
      gemm_rows_any(cur, sm.wS, oth, Kd, Nd, R);
      __syncthreads();
      const float* bl = a.arena + a.b_off[l];
      float* ysave = my_scratch + (size_t)(2 * l) * save_stride;
      float* zsave = ysave + save_stride;
      for (int idx = tid; idx < Nd * P; idx += kThreads) {
        const int n = idx / P, p = idx - n * P;
        float* row = oth + (size_t)n * kLd;
        const float av = tanhf(row[p] + __ldg(bl + n));
        row[p] = av;
        if (a.do_grad) { ysave[(size_t)n * kRows + p] = av; zsave[(size_t)n * kRows + p] = av; }
        if (J > 1) {
          const TanhF f(av);
          int c = 1;
          for (int i = 0; i < ndirs; ++i) {
            const int o = sg.dir_order[i];
            float z[4], y[4];
            for (int k = 0; k < o; ++k) z[k] = row[(c + k) * P + p];
            tanh_jet_fwd(f, z, o, y);
            for (int k = 0; k < o; ++k) {
              row[(c + k) * P + p] = y[k];
              if (a.do_grad) {
                ysave[(size_t)n * kRows + (c + k) * P + p] = y[k];
                zsave[(size_t)n * kRows + (c + k) * P + p] = z[k];
              }
            }
            c += o;
          }
        }
      }
      __syncthreads();
      float* t = cur; cur = oth; oth = t;
    }

    // ---- last layer: u[v][r] = sum_k Wl[v][k] * Y[k][r] (+ b_v on the value channel) ------------
    const int Kl = a.widths[L - 1];
    for (int idx = tid; idx < n_out * R; idx += kThreads) {
      const int v = idx / R, r = idx - v * R;
      float s = r < P ? a.arena[a.b_off[L - 1] + v] : 0.f;
      const float* wl = sm.wlS + v * kMaxW;
      for (int k = 0; k < Kl; ++k) s = fmaf(wl[k], cur[(size_t)k * kLd + r], s);
      sm.uS[idx] = s;
      sm.gvS[idx] = 0.f;
    }
    __syncthreads();

    // ---- virtual channels: V[v][m][g] = sum_{k,c} comb[m][k*J+c] * u[v][c][g*K+k] ----------------
    const float* V = sm.uS;
    if (!sg.identity) {
      const float* cm = a.comb + sg.comb_off;
      for (int idx = tid; idx < n_out * M * G; idx += kThreads) {
        const int g = idx % G, m = (idx / G) % M, v = idx / (G * M);
        float s = 0.f;
        for (int k = 0; k < K; ++k)
          for (int c = 0; c < J; ++c)
            s = fmaf(__ldg(cm + m * K * J + k * J + c), sm.uS[v * R + c * P + g * K + k], s);
        sm.vS[(v * M + m) * G + g] = s;
      }
      V = sm.vS;
      __syncthreads();
    }

    // ---- operator terms, residual, loss, adjoint seeds (one thread per row) ----------------------
    if (tid < g_valid) {
      const int g = tid;
      const long long row = g_first + g;
      for (int col = 0; col < ncols; ++col) {
        float val = 0.f;
        for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
          const tdb200_term tm = a.terms[t];
          float prod = tm.kind == 0 ? tm.coeff
                     : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                    : a.arena[a.n_net_params + tm.idx];
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = a.factors[fi];
            prod *= pow_i(V[(fc.var * M + fc.chan) * G + g], fc.ipow, fc.pow);
          }
          val += prod;
        }
        if (a.fields) a.fields[sg.field_off + row * ncols + col] = val;
        const float tgt = sg.tgt_off >= 0 ? __ldg(a.targets + sg.tgt_off + row * ncols + col) : 0.f;
        const float res = val - tgt;
        const int slot = sg.col_slot[col];
        const float rw = (a.row_weight && seg_i == 0) ? __ldg(a.row_weight + row) : 1.f;   // causal-loss weight (no grad)
        atomicAdd(&sm.lossS[slot], (double)rw * (double)res * (double)res);
        sm.rS[col * G + g] = jac ? (col == a.jac_col ? 1.f : 0.f)
                           : a.field_seed ? __ldg(a.field_seed + sg.field_off + row * ncols + col)
                                          : 2.f * __ldg(a.slot_scale + slot) * rw * res;
      }
      if (a.do_grad) {
        for (int col = 0; col < ncols; ++col) {
          const float seed = sm.rS[col * G + g];
          for (int t = sg.col_term_begin[col]; t < sg.col_term_end[col]; ++t) {
            const tdb200_term tm = a.terms[t];
            const float cf = tm.kind == 0 ? tm.coeff
                           : tm.kind == 1 ? __ldg(a.coeffs + tm.idx + row)
                                          : a.arena[a.n_net_params + tm.idx];
            float full = 1.f;
            for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
              const tdb200_factor fc = a.factors[fi];
              float part = seed * cf * dpow_i(V[(fc.var * M + fc.chan) * G + g], fc.ipow, fc.pow);
              for (int fj = tm.fac_begin; fj < tm.fac_end; ++fj) {
                if (fj == fi) continue;
                const tdb200_factor fo = a.factors[fj];
                part *= pow_i(V[(fo.var * M + fo.chan) * G + g], fo.ipow, fo.pow);
              }
              sm.gvS[(fc.var * M + fc.chan) * G + g] += part;
              full *= pow_i(V[(fc.var * M + fc.chan) * G + g], fc.ipow, fc.pow);
            }
            if (tm.kind == 2) atomicAdd(&sm.cgS[tm.idx], seed * full);
          }
        }
      }
    }
    __syncthreads();
    if (!a.do_grad) continue;

    // ---- adjoint of the virtual channels -> gu[v][r] ---------------------------------------------
    const float* GU = sm.gvS;
    if (!sg.identity) {
      const float* cm = a.comb + sg.comb_off;
      for (int idx = tid; idx < n_out * R; idx += kThreads) {
        const int v = idx / R, r = idx - v * R;
        const int c = r / P, p = r - c * P;
        const int g = p / K, k = p - g * K;
        float s = 0.f;
        for (int m = 0; m < M; ++m) s = fmaf(__ldg(cm + m * K * J + k * J + c), sm.gvS[(v * M + m) * G + g], s);
        sm.guS[idx] = s;
      }
      GU = sm.guS;
      __syncthreads();
    }

    // ---- backward of the last layer ----------------------------------------------------------------
    {
      float* dWl = my_grad + a.w_off[L - 1];
      for (int idx = tid; idx < n_out * Kl; idx += kThreads) {
        const int v = idx / Kl, k = idx - v * Kl;
        float s = 0.f;
        for (int r = 0; r < R; ++r) s = fmaf(GU[v * R + r], cur[(size_t)k * kLd + r], s);
        atomicAdd(dWl + idx, s);
      }
      if (tid < n_out) {
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += GU[tid * R + p];
        atomicAdd(my_grad + a.b_off[L - 1] + tid, s);
      }
      for (int idx = tid; idx < Kl * R; idx += kThreads) {
        const int k = idx / R, r = idx - k * R;
        float s = 0.f;
        for (int v = 0; v < n_out; ++v) s = fmaf(sm.wlS[v * kMaxW + k], GU[v * R + r], s);
        oth[(size_t)k * kLd + r] = s;
      }
    }
    __syncthreads();

    // ---- backward through the tanh layers ---------------------------------------------------------
    for (int t = L - 2; t >= 0; --t) {
      const int Nd = a.widths[t + 1];                       // width of this tanh layer's output
      const float* ysave = my_scratch + (size_t)(2 * t) * save_stride;
      const float* W0 = a.arena + a.w_off[0];
      // saved block of this layer ([a; z'; z''; ...], or just Y for layer 0) -> cur (free until Y_{t-1} is needed)
      {
        const float* blk = t == 0 ? ysave : ysave + save_stride;
        const int nr4 = (t == 0 ? P : R) / 4;
        for (int idx = tid; idx < Nd * nr4; idx += kThreads) {
          const int k = idx / nr4, r4 = idx - k * nr4;
          *reinterpret_cast<float4*>(cur + (size_t)k * kLd + r4 * 4) =
              *reinterpret_cast<const float4*>(blk + (size_t)k * kRows + r4 * 4);
        }
      }
      __syncthreads();
      // tanh-jet adjoint in place: oth holds gY, becomes gZ
      for (int idx = tid; idx < Nd * P; idx += kThreads) {
        const int n = idx / P, p = idx - n * P;
        float* row = oth + (size_t)n * kLd;
        const float* srow = cur + (size_t)n * kLd;
        const float av = srow[p];
        const TanhF f(av);
        float g0 = row[p] * f.f1;
        int c = 1;
        for (int i = 0; i < ndirs; ++i) {
          const int o = sg.dir_order[i];
          float z[4] = {0.f, 0.f, 0.f, 0.f}, gy[4], gz[4];
          if (t == 0) z[0] = dir_dot(W0 + n * d, sg.dir_vec[i], d);
          else for (int k = 0; k < o; ++k) z[k] = srow[(c + k) * P + p];
          for (int k = 0; k < o; ++k) gy[k] = row[(c + k) * P + p];
          g0 += tanh_jet_bwd(f, z, gy, o, gz);
          for (int k = 0; k < o; ++k) row[(c + k) * P + p] = gz[k];
          c += o;
        }
        row[p] = g0;
      }
      __syncthreads();
      // bias gradient: only the value channel carries the bias
      for (int n = tid; n < Nd; n += kThreads) {
        const float* row = oth + (size_t)n * kLd;
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += row[p];
        atomicAdd(my_grad + a.b_off[t] + n, s);
      }
      if (t == 0) {
        // dW0[n][ax] = sum_p gz0[n][p] x[p][ax] + sum_{dirs on ax} sum_p gz1[n][p]
        for (int idx = tid; idx < Nd * d; idx += kThreads) {
          const int n = idx / d, ax = idx - n * d;
          const float* row = oth + (size_t)n * kLd;
          float s = 0.f;
          for (int p = 0; p < P; ++p) s = fmaf(row[p], sm.xS[p * 4 + ax], s);
          int c = 1;
          for (int i = 0; i < ndirs; ++i) {
            const float vx = sg.dir_vec[i][ax];                // d z1 / d W0[n][ax] = component ax of the direction
            if (vx != 0.f) {
              float sd = 0.f;
              for (int p = 0; p < P; ++p) sd += row[c * P + p];
              s = fmaf(vx, sd, s);
            }
            c += sg.dir_order[i];
          }
          atomicAdd(my_grad + a.w_off[0] + idx, s);
        }
        __syncthreads();
        break;
      }
      const int Kd = a.widths[t];                           // input width of linear layer t
      // previous layer's output (all channels) back from scratch
      const float* yprev = my_scratch + (size_t)(2 * (t - 1)) * save_stride;
      for (int idx = tid; idx < Kd * (R / 4); idx += kThreads) {
        const int k = idx / (R / 4), r4 = idx - k * (R / 4);
        *reinterpret_cast<float4*>(cur + (size_t)k * kLd + r4 * 4) =
            *reinterpret_cast<const float4*>(yprev + (size_t)k * kRows + r4 * 4);
      }
      load_weight_image(sm.wS, a.img_b + (size_t)t * kMaxW * kWLd, Nd);   // W_t as [n][k]: reduction over n
      __syncthreads();
      gemm_wgrad_any(oth, cur, my_grad + a.w_off[t], Nd, Kd, R);
      __syncthreads();
      gemm_rows_any(oth, sm.wS, cur, Nd, Kd, R);            // gY_{t-1}[k][r] = sum_n W[n][k] gZ[n][r]
      __syncthreads();
      float* tmp = cur; cur = oth; oth = tmp;
    }
    if (jac) {                                              // coefficient-parameter entries of this row
      __syncthreads();
      if (tid < a.n_cparams) { my_grad[a.n_net_params + tid] = sm.cgS[tid]; sm.cgS[tid] = 0.f; }
    }
  }

  // ---- flush per-CTA scalars ------------------------------------------------------------------------
  __syncthreads();
  if (!jac && tid < a.n_slots) a.part_loss[(size_t)blockIdx.x * a.n_slots + tid] = sm.lossS[tid];
  if (a.do_grad && !jac && tid < a.n_cparams) my_grad[a.n_net_params + tid] = sm.cgS[tid];
}

cudaError_t launch_jet_simt(const JetArgs& a, int grid, cudaStream_t s) {
  const size_t smem = jet_simt_smem_bytes(a.wmax);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(jet_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  jet_simt_kernel<<<grid, kThreads, smem, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// deterministic cross-CTA reduction + loss assembly
// ------------------------------------------------------------------------------------------------
// Block = 64 consecutive parameters x 8 row groups: thread (p, g) adds the rows g, g + 8, ... of its parameter (loads of
// 8 rows in flight), the 8 group sums are combined in a fixed order -> bit-reproducible, and 8x the parallelism of one
// thread per parameter (BASELINE config 1: 20 -> ~6 us of a 147 us step).
constexpr int kRedParams = 64, kRedGroups = 8;
__global__ void __launch_bounds__(kRedParams * kRedGroups)
reduce_partials_kernel(const float* __restrict__ part_grad, int n_grad_rows, const double* __restrict__ part_loss,
                       int n_loss_rows, int n_params, int n_params_pad, int n_slots,
                       const double* __restrict__ slot_lambda, const double* __restrict__ slot_len,
                       float* __restrict__ out) {
  pdl_wait();
  __shared__ float gsum[kRedGroups][kRedParams];
  const int pl = threadIdx.x % kRedParams, g = threadIdx.x / kRedParams;
  const int i = blockIdx.x * kRedParams + pl;
  float s = 0.f;
  if (i < n_params) {
    int c = g;
    for (; c + 7 * kRedGroups < n_grad_rows; c += 8 * kRedGroups) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(part_grad + (size_t)(c + k * kRedGroups) * n_params_pad + i);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
    for (; c < n_grad_rows; c += kRedGroups) s += __ldg(part_grad + (size_t)c * n_params_pad + i);
  }
  gsum[g][pl] = s;
  __syncthreads();
  if (g == 0 && i < n_params) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kRedGroups; ++k) t += gsum[k][pl];
    out[2 + n_slots + i] = t;
  }
  if (blockIdx.x == gridDim.x - 1) {
    // loss terms: one warp per slot (lanes stride the rows, then a fixed shuffle tree); warp 0 assembles the loss
    __shared__ double mse_s[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int sl = warp; sl < n_slots; sl += n_warps) {
      double acc = 0.0;
      for (int c = lane; c < n_loss_rows; c += 32) acc += part_loss[(size_t)c * n_slots + sl];
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) { mse_s[sl] = acc / slot_len[sl]; out[2 + sl] = (float)mse_s[sl]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double loss = 0.0, lossn = 0.0;
      for (int sl = 0; sl < n_slots; ++sl) { loss += slot_lambda[sl] * mse_s[sl]; lossn += mse_s[sl]; }
      out[0] = (float)loss;
      out[1] = (float)lossn;
    }
  }
}

cudaError_t launch_reduce_partials(const float* part_grad, int n_grad_rows, const double* part_loss,
                                   int n_loss_rows, int n_params, int n_params_pad, int n_slots,
                                   const double* slot_lambda, const double* slot_len, float* out, cudaStream_t s) {
  const int blocks = max(1, (n_params + kRedParams - 1) / kRedParams);
  return launch_pdl(reduce_partials_kernel, dim3(blocks), dim3(kRedParams * kRedGroups), 0, s, part_grad, n_grad_rows, part_loss, n_loss_rows, n_params,
                                                                   n_params_pad, n_slots, slot_lambda, slot_len, out);
  return cudaGetLastError();
}

}  // namespace tdb
