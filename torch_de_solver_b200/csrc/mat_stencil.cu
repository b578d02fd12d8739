// mat-mode residual + adjoint stencil kernels (sm_100a), bandwidth bound.
//
// Replaces Derivative_mat (tedeous/derivative.py:135-323: ~10 torch.roll / index_put passes per derivative and a
// torch.unique per call), the mat branches of Operator/Bounds (tedeous/eval.py:143-193, 298-302, 320-326), the
// MSE of losses.py:84-135 and the autograd backward through all of it (optimizers/closure.py:60) by
//   1. one tiled kernel: load u (+ halo) to shared memory, evaluate every derivative field F_q = D_a^k u_v
//      as a banded 1-D stencil (edge rows have their own coefficients - SURVEY Appendix D), the operator
//      terms and the residual on tile + halo, reduce the loss, form the field adjoints A_q in shared memory
//      and apply the transposed stencils -> d loss / d u for the tile.  HBM traffic: read u, read the
//      coefficient tensors, write the gradient (12 B / cell for Poisson) + halo re-reads served by L2.
//   2. a small kernel for the boundary rows (gather, residual, scatter-add of the adjoint),
//   3. a finalize kernel (ordered reduction of per-CTA loss partials, loss assembly).
#include <string>
#include <vector>
#include "common.cuh"

namespace tdb {

constexpr int kMatTY = 32, kMatTX = 64, kMatThreads = 256;
constexpr int kMatMaxFields = 12;
constexpr int kMatMaxVar = 4;
constexpr int kMatMaxHalo = 8;
constexpr int kMatMaxTaps = 64;
constexpr int kMatMaxForcing = 16;

struct MatArgs {
  int n_eq, n_var, n0, n1, n_fields;
  int hy, hx;                              // halo of the derivative fields along axis 0 / 1
  tdb200_mat_field fld[kMatMaxFields];
  int eq_term_begin[TDB200_MAX_COLS], eq_term_end[TDB200_MAX_COLS];
  float eq_scale[TDB200_MAX_COLS];         // lambda_eq / (n0*n1)
  const float* band;
  const tdb200_term* terms;
  const tdb200_factor* factors;
  const float* coeffs;
  const float* u;
  float* grad;
  float* op_out;                           // optional [n0*n1][n_eq]
  double* part_loss;                       // [n_ctas][n_eq]
  int tiles_x, tiles_y;
  // fast path (every equation = constant-coefficient linear terms + forcing): composite interior taps
  int linear;                              // 1: tables below are valid
  int edge_y, edge_x;                      // rows / columns with special (one-sided) coefficients at each end
  int tap_begin[TDB200_MAX_COLS + 1];      // taps of equation e: [tap_begin[e], tap_begin[e+1])
  short tap_var[kMatMaxTaps], tap_axis[kMatMaxTaps], tap_m[kMatMaxTaps];
  float tap_w[kMatMaxTaps];
  int n_lin;                               // linear terms (all equations), for edge cells of the fast path
  short lin_eq[kMatMaxTaps], lin_q[kMatMaxTaps];
  float lin_c[kMatMaxTaps];
  int lin1;                                // 1: single equation, single field, <= 16 taps -> register-tap kernel path
  int frc_begin[TDB200_MAX_COLS + 1];      // forcing terms of equation e
  float frc_const[kMatMaxForcing];
  long long frc_buf[kMatMaxForcing];       // coefficient buffer offset, -1: constant only
};

__device__ __forceinline__ float band_coef(const float* __restrict__ band, const tdb200_mat_field& f, int n, int i, int m) {
  const int w = 2 * f.half_width + 1;
  const float* base = band + f.coef_off;
  if (i < f.n_edge) return __ldg(base + w + i * w + m + f.half_width);
  if (i >= n - f.n_edge) return __ldg(base + w + f.n_edge * w + (n - 1 - i) * w + m + f.half_width);
  return __ldg(base + m + f.half_width);
}

// value of derivative field f at region cell (ly, lx) of the shared u tile
__device__ __forceinline__ float field_value(const MatArgs& a, const tdb200_mat_field& f, const float* __restrict__ us,
                                             int pitch, int plane, int ly, int lx, int gy, int gx) {
  const float* up = us + f.var * plane;
  if (f.order == 0) return up[ly * pitch + lx];
  float s = 0.f;
  const int b = f.half_width;
  if (f.axis == 0) {
    for (int m = -b; m <= b; ++m) {
      const int yy = gy + m;
      if (yy < 0 || yy >= a.n0) continue;
      s = fmaf(band_coef(a.band, f, a.n0, gy, m), up[(ly + m) * pitch + lx], s);
    }
  } else {
    for (int m = -b; m <= b; ++m) {
      const int xx = gx + m;
      if (xx < 0 || xx >= a.n1) continue;
      s = fmaf(band_coef(a.band, f, a.n1, gx, m), up[ly * pitch + lx + m], s);
    }
  }
  return s;
}

// ---- register-tap path: one linear constant-coefficient equation on one field -------------------------
// Interior cells use a composite tap list held in registers; the few cells whose stencil rows are special
// (one-sided rows near the domain edge) evaluate the banded operators directly.
template <int NT>
__device__ __forceinline__ void mat_lin1_path(const MatArgs& a, float* us, float* ss, double (*red)[TDB200_MAX_COLS],
                                              int ty0, int tx0) {
  const int hy = a.hy, hx = a.hx;
  const int ux = kMatTX + 4 * hx;
  const int ry = kMatTY + 2 * hy, rx = kMatTX + 2 * hx;
  const int tid = threadIdx.x;
  const int nt = a.tap_begin[1];
  float tw[NT];
  int tou[NT], tos[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const bool on = t < nt;
    tw[t] = on ? a.tap_w[t] : 0.f;
    const int m = on ? a.tap_m[t] : 0;
    const bool ax0 = on && a.tap_axis[t] == 0;
    tou[t] = ax0 ? m * ux : m;
    tos[t] = ax0 ? -m * rx : -m;
  }
  const float scale2 = 2.f * a.eq_scale[0];
  const int zy = a.edge_y, zx = a.edge_x;                 // rows / columns with their own coefficients
  float lacc = 0.f;
  // phase 2: residual seeds on tile + halo
  {
    int ly = tid / rx, lx = tid - ly * rx;
    const int dly = kMatThreads / rx, dlx = kMatThreads - dly * rx;
    for (; ly < ry; ly += dly, lx += dlx) {
      if (lx >= rx) { lx -= rx; ++ly; if (ly >= ry) break; }
      const int gy = ty0 - hy + ly, gx = tx0 - hx + lx;
      float seed = 0.f;
      if (gy >= 0 && gy < a.n0 && gx >= 0 && gx < a.n1) {
        const size_t cell = (size_t)gy * a.n1 + gx;
        float res = 0.f;
        for (int t = a.frc_begin[0]; t < a.frc_begin[1]; ++t)
          res += a.frc_buf[t] >= 0 ? __ldg(a.coeffs + a.frc_buf[t] + cell) : a.frc_const[t];
        const float* uc = us + (ly + hy) * ux + lx + hx;
        if (gy >= zy && gy < a.n0 - zy && gx >= zx && gx < a.n1 - zx) {
#pragma unroll
          for (int t = 0; t < NT; ++t) res = fmaf(tw[t], uc[tou[t]], res);
        } else {
          for (int t = 0; t < a.n_lin; ++t)
            res = fmaf(a.lin_c[t], field_value(a, a.fld[a.lin_q[t]], us, ux, 0, ly + hy, lx + hx, gy, gx), res);
        }
        if (ly >= hy && ly < hy + kMatTY && lx >= hx && lx < hx + kMatTX) {
          lacc += res * res;
          if (a.op_out) a.op_out[cell] = res;
        }
        seed = scale2 * res;
      }
      ss[ly * rx + lx] = seed;
    }
  }
  {
    double v = (double)lacc;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5][0] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kMatThreads / 32; ++w) s += red[w][0];
    a.part_loss[(size_t)(blockIdx.y * gridDim.x + blockIdx.x)] = s;
  }
  if (!a.grad) return;
  // phase 3: transposed stencil -> gradient of the tile (kMatTX == 64: lx = tid & 63)
  const int zy3 = zy + hy, zx3 = zx + hx;                  // cells that gather from a special row
  for (int cy = tid >> 6; cy < kMatTY; cy += kMatThreads >> 6) {
    const int cx = tid & 63;
    const int gy = ty0 + cy, gx = tx0 + cx;
    if (gy >= a.n0 || gx >= a.n1) continue;
    const float* sc = ss + (cy + hy) * rx + cx + hx;
    float g = 0.f;
    if (gy >= zy3 && gy < a.n0 - zy3 && gx >= zx3 && gx < a.n1 - zx3) {
#pragma unroll
      for (int t = 0; t < NT; ++t) g = fmaf(tw[t], sc[tos[t]], g);
    } else {
      for (int t = 0; t < a.n_lin; ++t) {
        const tdb200_mat_field& f = a.fld[a.lin_q[t]];
        float s = 0.f;
        if (f.order == 0) s = sc[0];
        else if (f.axis == 0) {
          for (int m = -f.half_width; m <= f.half_width; ++m) {
            const int yy = gy + m;
            if (yy >= 0 && yy < a.n0) s = fmaf(band_coef(a.band, f, a.n0, yy, -m), sc[m * rx], s);
          }
        } else {
          for (int m = -f.half_width; m <= f.half_width; ++m) {
            const int xx = gx + m;
            if (xx >= 0 && xx < a.n1) s = fmaf(band_coef(a.band, f, a.n1, xx, -m), sc[m], s);
          }
        }
        g = fmaf(a.lin_c[t], s, g);
      }
    }
    a.grad[(size_t)gy * a.n1 + gx] = g;
  }
}

__global__ void __launch_bounds__(kMatThreads) mat_residual_adjoint_kernel(const MatArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int hy = a.hy, hx = a.hx;
  const int uy = kMatTY + 4 * hy, ux = kMatTX + 4 * hx;       // u region (tile + 2 halos)
  const int ry = kMatTY + 2 * hy, rx = kMatTX + 2 * hx;       // residual / adjoint region (tile + halo)
  const int uplane = uy * ux, rplane = ry * rx;
  float* us = sm;                                            // [n_var][uy][ux]
  float* as = us + a.n_var * uplane;                         // [n_fields][ry][rx]
  __shared__ double red[kMatThreads / 32][TDB200_MAX_COLS];

  const int ty0 = blockIdx.y * kMatTY, tx0 = blockIdx.x * kMatTX;
  const int tid = threadIdx.x;
  const size_t N = (size_t)a.n0 * a.n1;

  // ---- phase 1: u tile + 2 halos -> smem (zero outside the domain) --------------------------------
  for (int v = 0; v < a.n_var; ++v) {
    int ly = tid / ux, lx = tid - ly * ux;
    const int dly = kMatThreads / ux, dlx = kMatThreads - dly * ux;
    for (; ly < uy; ly += dly, lx += dlx) {
      if (lx >= ux) { lx -= ux; ++ly; if (ly >= uy) break; }
      const int gy = ty0 - 2 * hy + ly, gx = tx0 - 2 * hx + lx;
      float val = 0.f;
      if (gy >= 0 && gy < a.n0 && gx >= 0 && gx < a.n1) val = __ldg(a.u + (size_t)v * N + (size_t)gy * a.n1 + gx);
      us[v * uplane + ly * ux + lx] = val;
    }
  }
  __syncthreads();

  if (a.lin1) {
    const int nt = a.tap_begin[1];
    if (nt <= 4) mat_lin1_path<4>(a, us, as, red, ty0, tx0);
    else if (nt <= 6) mat_lin1_path<6>(a, us, as, red, ty0, tx0);
    else if (nt <= 8) mat_lin1_path<8>(a, us, as, red, ty0, tx0);
    else if (nt <= 12) mat_lin1_path<12>(a, us, as, red, ty0, tx0);
    else mat_lin1_path<16>(a, us, as, red, ty0, tx0);
    return;
  }

  // ---- fast path: interior tile of a linear constant-coefficient operator ---------------------------
  // (uniform per CTA) the residual is one composite stencil, the gradient its transpose applied to the seeds
  const bool fast = a.linear && ty0 - 2 * hy >= a.edge_y && ty0 + kMatTY + 2 * hy <= a.n0 - a.edge_y &&
                    tx0 - 2 * hx >= a.edge_x && tx0 + kMatTX + 2 * hx <= a.n1 - a.edge_x;
  if (fast) {
    float lacc[TDB200_MAX_COLS];
#pragma unroll
    for (int e = 0; e < TDB200_MAX_COLS; ++e) lacc[e] = 0.f;
    float* ss = as;                                          // [n_eq][ry][rx] residual seeds
    for (int idx = tid; idx < rplane; idx += kMatThreads) {
      const int ly = idx / rx, lx = idx - ly * rx;
      const int gy = ty0 - hy + ly, gx = tx0 - hx + lx;
      const size_t cell = (size_t)gy * a.n1 + gx;
      const bool core = ly >= hy && ly < hy + kMatTY && lx >= hx && lx < hx + kMatTX;
      const float* uc = us + (ly + hy) * ux + lx + hx;
      for (int e = 0; e < a.n_eq; ++e) {
        float res = 0.f;
        for (int t = a.frc_begin[e]; t < a.frc_begin[e + 1]; ++t)
          res += a.frc_buf[t] >= 0 ? __ldg(a.coeffs + a.frc_buf[t] + cell) : a.frc_const[t];
        for (int t = a.tap_begin[e]; t < a.tap_begin[e + 1]; ++t) {
          const int off = a.tap_var[t] * uplane + (a.tap_axis[t] == 0 ? a.tap_m[t] * ux : a.tap_m[t]);
          res = fmaf(a.tap_w[t], uc[off], res);
        }
        if (core) {
          lacc[e] += res * res;
          if (a.op_out) a.op_out[cell * a.n_eq + e] = res;
        }
        ss[e * rplane + idx] = 2.f * a.eq_scale[e] * res;
      }
    }
    for (int e = 0; e < a.n_eq; ++e) {
      double v = (double)lacc[e];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) red[tid >> 5][e] = v;
    }
    __syncthreads();
    if (tid < a.n_eq) {
      double s = 0.0;
      for (int w = 0; w < kMatThreads / 32; ++w) s += red[w][tid];
      a.part_loss[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * a.n_eq + tid] = s;
    }
    if (!a.grad) return;
    for (int idx = tid; idx < kMatTY * kMatTX; idx += kMatThreads) {
      const int cy = idx / kMatTX, cx = idx - cy * kMatTX;
      const float* sc = ss + (cy + hy) * rx + cx + hx;
      float g[kMatMaxVar];
#pragma unroll
      for (int v = 0; v < kMatMaxVar; ++v) g[v] = 0.f;
      for (int e = 0; e < a.n_eq; ++e)
        for (int t = a.tap_begin[e]; t < a.tap_begin[e + 1]; ++t) {
          const int off = e * rplane - (a.tap_axis[t] == 0 ? a.tap_m[t] * rx : a.tap_m[t]);
          const float c = a.tap_w[t] * sc[off];
#pragma unroll
          for (int v = 0; v < kMatMaxVar; ++v) if (v == a.tap_var[t]) g[v] += c;
        }
      const size_t cell = (size_t)(ty0 + cy) * a.n1 + tx0 + cx;
      for (int v = 0; v < a.n_var; ++v) a.grad[(size_t)v * N + cell] = g[v];
    }
    return;
  }

  // ---- phase 2: fields, residual, loss, field adjoints on tile + halo -------------------------------
  float loss_acc[TDB200_MAX_COLS];
#pragma unroll
  for (int e = 0; e < TDB200_MAX_COLS; ++e) loss_acc[e] = 0.f;
  for (int idx = tid; idx < rplane; idx += kMatThreads) {
    const int ly = idx / rx, lx = idx - ly * rx;
    const int gy = ty0 - hy + ly, gx = tx0 - hx + lx;
    float F[kMatMaxFields], A[kMatMaxFields];
    const bool inside = gy >= 0 && gy < a.n0 && gx >= 0 && gx < a.n1;
#pragma unroll
    for (int q = 0; q < kMatMaxFields; ++q) A[q] = 0.f;
    if (inside) {
#pragma unroll
      for (int q = 0; q < kMatMaxFields; ++q)
        if (q < a.n_fields) F[q] = field_value(a, a.fld[q], us, ux, uplane, ly + hy, lx + hx, gy, gx);
      const size_t cell = (size_t)gy * a.n1 + gx;
      const bool core = ly >= hy && ly < hy + kMatTY && lx >= hx && lx < hx + kMatTX;
      for (int e = 0; e < a.n_eq; ++e) {
        float res = 0.f;
        for (int t = a.eq_term_begin[e]; t < a.eq_term_end[e]; ++t) {
          const tdb200_term tm = a.terms[t];
          float prod = tm.kind == 1 ? __ldg(a.coeffs + tm.idx + cell) : tm.coeff;
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = a.factors[fi];
            float x = 0.f;
#pragma unroll
            for (int q = 0; q < kMatMaxFields; ++q) if (q == fc.chan) x = F[q];
            prod *= pow_i(x, fc.ipow, fc.pow);
          }
          res += prod;
        }
        if (core) {
          loss_acc[e] += res * res;
          if (a.op_out) a.op_out[cell * a.n_eq + e] = res;
        }
        const float seed = 2.f * a.eq_scale[e] * res;
        for (int t = a.eq_term_begin[e]; t < a.eq_term_end[e]; ++t) {
          const tdb200_term tm = a.terms[t];
          const float cf = tm.kind == 1 ? __ldg(a.coeffs + tm.idx + cell) : tm.coeff;
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = a.factors[fi];
            float x = 0.f;
#pragma unroll
            for (int q = 0; q < kMatMaxFields; ++q) if (q == fc.chan) x = F[q];
            float part = seed * cf * dpow_i(x, fc.ipow, fc.pow);
            for (int fj = tm.fac_begin; fj < tm.fac_end; ++fj) {
              if (fj == fi) continue;
              const tdb200_factor fo = a.factors[fj];
              float xo = 0.f;
#pragma unroll
              for (int q = 0; q < kMatMaxFields; ++q) if (q == fo.chan) xo = F[q];
              part *= pow_i(xo, fo.ipow, fo.pow);
            }
#pragma unroll
            for (int q = 0; q < kMatMaxFields; ++q) if (q == fc.chan) A[q] += part;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kMatMaxFields; ++q)
      if (q < a.n_fields) as[q * rplane + idx] = A[q];
  }
  // loss partial of this CTA (fixed order: warp shuffle tree, then warps in order)
  for (int e = 0; e < a.n_eq; ++e) {
    double v = (double)loss_acc[e];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5][e] = v;
  }
  __syncthreads();
  if (tid < a.n_eq) {
    double s = 0.0;
    for (int w = 0; w < kMatThreads / 32; ++w) s += red[w][tid];
    a.part_loss[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * a.n_eq + tid] = s;
  }
  if (!a.grad) return;

  // ---- phase 3: transposed stencils -> gradient of the tile ----------------------------------------
  for (int idx = tid; idx < kMatTY * kMatTX; idx += kMatThreads) {
    const int cy = idx / kMatTX, cx = idx - cy * kMatTX;
    const int gy = ty0 + cy, gx = tx0 + cx;
    if (gy >= a.n0 || gx >= a.n1) continue;
    const int ly = cy + hy, lx = cx + hx;                    // position in the adjoint region
    float g[kMatMaxVar];
#pragma unroll
    for (int v = 0; v < kMatMaxVar; ++v) g[v] = 0.f;
    for (int q = 0; q < a.n_fields; ++q) {
      const tdb200_mat_field& f = a.fld[q];
      const float* ap = as + q * rplane;
      float s = 0.f;
      if (f.order == 0) {
        s = ap[ly * rx + lx];
      } else if (f.axis == 0) {
        for (int m = -f.half_width; m <= f.half_width; ++m) {
          const int yy = gy + m;                             // row of D that touches column gy
          if (yy < 0 || yy >= a.n0) continue;
          s = fmaf(band_coef(a.band, f, a.n0, yy, -m), ap[(ly + m) * rx + lx], s);
        }
      } else {
        for (int m = -f.half_width; m <= f.half_width; ++m) {
          const int xx = gx + m;
          if (xx < 0 || xx >= a.n1) continue;
          s = fmaf(band_coef(a.band, f, a.n1, xx, -m), ap[ly * rx + lx + m], s);
        }
      }
#pragma unroll
      for (int v = 0; v < kMatMaxVar; ++v) if (v == f.var) g[v] += s;
    }
    for (int v = 0; v < a.n_var; ++v) a.grad[(size_t)v * N + (size_t)gy * a.n1 + gx] = g[v];
  }
}

// ---- boundary rows -----------------------------------------------------------------------------------
struct MatBcArgs {
  int n_var, n0, n1, n_fields, n_eq;
  tdb200_mat_field fld[kMatMaxFields];
  const float* band;
  const tdb200_term* terms;
  const tdb200_factor* factors;
  const float* coeffs;
  const tdb200_mat_bc* bcs;
  int n_bcs;
  const long long* bc_row_begin;           // [n_bcs + 1] prefix of n_rows
  const int* cells;
  const float* targets;
  const float* slot_scale;                 // boundary slots: lambda / max_len
  const float* u;
  float* grad;
  float* bval_out;                         // optional, per row
  double* slot_sum;                        // [n_bc_slots]
};

__device__ float global_field(const MatBcArgs& a, const tdb200_mat_field& f, int cell) {
  const size_t N = (size_t)a.n0 * a.n1;
  const float* up = a.u + (size_t)f.var * N;
  if (f.order == 0) return up[cell];
  const int gy = cell / a.n1, gx = cell - gy * a.n1;
  float s = 0.f;
  for (int m = -f.half_width; m <= f.half_width; ++m) {
    if (f.axis == 0) {
      const int yy = gy + m;
      if (yy < 0 || yy >= a.n0) continue;
      s = fmaf(band_coef(a.band, f, a.n0, gy, m), up[(size_t)yy * a.n1 + gx], s);
    } else {
      const int xx = gx + m;
      if (xx < 0 || xx >= a.n1) continue;
      s = fmaf(band_coef(a.band, f, a.n1, gx, m), up[(size_t)gy * a.n1 + xx], s);
    }
  }
  return s;
}

__device__ void scatter_field_adjoint(const MatBcArgs& a, const tdb200_mat_field& f, int cell, float g) {
  const size_t N = (size_t)a.n0 * a.n1;
  float* gp = a.grad + (size_t)f.var * N;
  if (f.order == 0) { atomicAdd(gp + cell, g); return; }
  const int gy = cell / a.n1, gx = cell - gy * a.n1;
  for (int m = -f.half_width; m <= f.half_width; ++m) {
    if (f.axis == 0) {
      const int yy = gy + m;
      if (yy < 0 || yy >= a.n0) continue;
      atomicAdd(gp + (size_t)yy * a.n1 + gx, g * band_coef(a.band, f, a.n0, gy, m));
    } else {
      const int xx = gx + m;
      if (xx < 0 || xx >= a.n1) continue;
      atomicAdd(gp + (size_t)gy * a.n1 + xx, g * band_coef(a.band, f, a.n1, gx, m));
    }
  }
}

__global__ void mat_bc_kernel(const MatBcArgs a) {
  const long long total = a.bc_row_begin[a.n_bcs];
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < total;
       row += (long long)gridDim.x * blockDim.x) {
    int bi = 0;
    while (row >= a.bc_row_begin[bi + 1]) ++bi;
    const tdb200_mat_bc bc = a.bcs[bi];
    const long long r = row - a.bc_row_begin[bi];
    float val = 0.f;
    for (int k = 0; k < bc.K; ++k) {
      const int cell = a.cells[bc.cell_off + r * bc.K + k];
      float v = 0.f;
      if (bc.term_begin == bc.term_end) {
        v = a.u[(size_t)bc.var * a.n0 * a.n1 + cell];
      } else {
        for (int t = bc.term_begin; t < bc.term_end; ++t) {
          const tdb200_term tm = a.terms[t];
          float prod = tm.kind == 1 ? a.coeffs[tm.idx + cell] : tm.coeff;
          for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
            const tdb200_factor fc = a.factors[fi];
            prod *= pow_i(global_field(a, a.fld[fc.chan], cell), fc.ipow, fc.pow);
          }
          v += prod;
        }
      }
      val += bc.sign[k] * v;
    }
    if (a.bval_out) a.bval_out[row] = val;
    const float res = val - a.targets[bc.tgt_off + r];
    atomicAdd(a.slot_sum + bc.slot, (double)res * (double)res);
    if (!a.grad) continue;
    const float seed = 2.f * a.slot_scale[bc.slot] * res;
    for (int k = 0; k < bc.K; ++k) {
      const int cell = a.cells[bc.cell_off + r * bc.K + k];
      const float sk = seed * bc.sign[k];
      if (bc.term_begin == bc.term_end) {
        atomicAdd(a.grad + (size_t)bc.var * a.n0 * a.n1 + cell, sk);
        continue;
      }
      for (int t = bc.term_begin; t < bc.term_end; ++t) {
        const tdb200_term tm = a.terms[t];
        const float cf = tm.kind == 1 ? a.coeffs[tm.idx + cell] : tm.coeff;
        for (int fi = tm.fac_begin; fi < tm.fac_end; ++fi) {
          const tdb200_factor fc = a.factors[fi];
          float part = sk * cf * dpow_i(global_field(a, a.fld[fc.chan], cell), fc.ipow, fc.pow);
          for (int fj = tm.fac_begin; fj < tm.fac_end; ++fj) {
            if (fj == fi) continue;
            const tdb200_factor fo = a.factors[fj];
            part *= pow_i(global_field(a, a.fld[fo.chan], cell), fo.ipow, fo.pow);
          }
          scatter_field_adjoint(a, a.fld[fc.chan], cell, part);
        }
      }
    }
  }
}

__global__ void mat_finalize_kernel(const double* __restrict__ part_loss, int n_ctas, int n_eq, double n_cells,
                                    const double* __restrict__ bc_sum, int n_bc_slots,
                                    const double* __restrict__ slot_lambda, const double* __restrict__ slot_len,
                                    float* __restrict__ out) {
  __shared__ double sh[256];
  __shared__ double mse[32];
  for (int e = 0; e < n_eq; ++e) {
    double s = 0.0;
    for (int c = threadIdx.x; c < n_ctas; c += blockDim.x) s += part_loss[(size_t)c * n_eq + e];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) mse[e] = sh[0] / n_cells;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double loss = 0.0, lossn = 0.0;
    for (int e = 0; e < n_eq; ++e) { out[2 + e] = (float)mse[e]; loss += slot_lambda[e] * mse[e]; lossn += mse[e]; }
    for (int s = 0; s < n_bc_slots; ++s) {
      const double m = bc_sum[s] / slot_len[n_eq + s];
      out[2 + n_eq + s] = (float)m;
      loss += slot_lambda[n_eq + s] * m;
      lossn += m;
    }
    out[0] = (float)loss;
    out[1] = (float)lossn;
  }
}

}  // namespace tdb

// ------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_mat_err;
}
extern "C" const char* tdb200_last_error(void);
extern "C" void tdb200_set_error_(const char* msg);

struct tdb200_mat_plan {
  tdb200_mat_desc desc{};
  int device = 0;
  tdb::MatArgs args{};
  tdb::MatBcArgs bc{};
  std::vector<tdb200_mat_field> fields;
  float* d_band = nullptr;
  tdb200_term* d_terms = nullptr;
  tdb200_factor* d_factors = nullptr;
  tdb200_mat_bc* d_bcs = nullptr;
  long long* d_bc_row_begin = nullptr;
  float* d_slot_scale = nullptr;
  double* d_slot_lambda = nullptr;
  double* d_slot_len = nullptr;
  double* d_part_loss = nullptr;
  double* d_bc_sum = nullptr;
  int n_ctas = 0;
  int n_bc_slots = 0;
  int n_slots = 0;
  long long n_bc_rows = 0;
  size_t smem = 0;
  bool bcs_set = false;
  std::vector<double> lambda_eq;
};

#define MCU(call)                                                          \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      tdb200_set_error_((std::string(#call) + ": " + cudaGetErrorString(e__)).c_str()); \
      return TDB200_ERR_CUDA;                                              \
    }                                                                      \
  } while (0)

static int mat_invalid(const char* msg) {
  tdb200_set_error_(msg);
  return TDB200_ERR_INVALID;
}

extern "C" {

int tdb200_mat_plan_create(const tdb200_mat_desc* desc, const tdb200_mat_field* fields, int32_t n_band,
                           const float* band, const int32_t* eq_term_begin, const int32_t* eq_term_end,
                           int32_t n_terms, const tdb200_term* terms, int32_t n_factors,
                           const tdb200_factor* factors, int32_t device, tdb200_mat_plan** out) {
  if (!desc || !fields || !band || !eq_term_begin || !eq_term_end || !out) return mat_invalid("null argument");
  if (desc->n_eq < 1 || desc->n_eq > TDB200_MAX_COLS) return mat_invalid("n_eq out of range");
  if (desc->n_var < 1 || desc->n_var > tdb::kMatMaxVar) return mat_invalid("n_var out of range (1..4)");
  if (desc->n_fields < desc->n_var || desc->n_fields > tdb::kMatMaxFields) return mat_invalid("n_fields out of range");
  if (desc->n0 < 1 || desc->n1 < 1) return mat_invalid("empty grid");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    tdb200_set_error_("no CUDA device");
    return TDB200_ERR_NO_DEVICE;
  }
  MCU(cudaSetDevice(device));
  auto* p = new tdb200_mat_plan();
  p->desc = *desc;
  p->device = device;
  tdb::MatArgs& a = p->args;
  a.n_eq = desc->n_eq; a.n_var = desc->n_var; a.n0 = desc->n0; a.n1 = desc->n1; a.n_fields = desc->n_fields;
  int hy = 0, hx = 0;
  for (int q = 0; q < desc->n_fields; ++q) {
    a.fld[q] = fields[q];
    p->bc.fld[q] = fields[q];
    if (fields[q].order > 0) {
      if (fields[q].axis == 0) hy = fields[q].half_width > hy ? fields[q].half_width : hy;
      else hx = fields[q].half_width > hx ? fields[q].half_width : hx;
    }
    if (fields[q].var < 0 || fields[q].var >= desc->n_var || fields[q].axis < 0 || fields[q].axis > 1) {
      delete p; return mat_invalid("bad field");
    }
  }
  if (hy > tdb::kMatMaxHalo || hx > tdb::kMatMaxHalo) { delete p; return mat_invalid("stencil reach too large"); }
  a.hy = hy; a.hx = hx;
  for (int e = 0; e < desc->n_eq; ++e) { a.eq_term_begin[e] = eq_term_begin[e]; a.eq_term_end[e] = eq_term_end[e]; }
  {  // fast-path analysis: constant-coefficient linear terms + forcing only
    bool linear = true;
    int n_taps = 0, n_frc = 0, ey = 0, ex = 0, n_lin = 0;
    for (int e = 0; e < desc->n_eq && linear; ++e) {
      a.tap_begin[e] = n_taps;
      a.frc_begin[e] = n_frc;
      for (int t = eq_term_begin[e]; t < eq_term_end[e] && linear; ++t) {
        const tdb200_term& tm = terms[t];
        int live = 0, fq = -1;
        for (int f = tm.fac_begin; f < tm.fac_end; ++f) {
          if (factors[f].ipow == 0) continue;
          ++live;
          fq = factors[f].chan;
          if (factors[f].ipow != 1) linear = false;
        }
        if (live == 0) {
          if (n_frc >= tdb::kMatMaxForcing || tm.kind == 2) { linear = false; break; }
          a.frc_const[n_frc] = tm.kind == 0 ? tm.coeff : 0.f;
          a.frc_buf[n_frc] = tm.kind == 1 ? tm.idx : -1;
          ++n_frc;
        } else if (live == 1 && tm.kind == 0 && linear) {
          const tdb200_mat_field& f = fields[fq];
          if (n_lin < tdb::kMatMaxTaps) { a.lin_eq[n_lin] = (short)e; a.lin_q[n_lin] = (short)fq; a.lin_c[n_lin] = tm.coeff; }
          ++n_lin;
          if (f.order == 0) {
            if (n_taps >= tdb::kMatMaxTaps) { linear = false; break; }
            a.tap_var[n_taps] = (short)f.var; a.tap_axis[n_taps] = 1; a.tap_m[n_taps] = 0; a.tap_w[n_taps] = tm.coeff;
            ++n_taps;
          } else {
            const float* interior = band + f.coef_off;
            for (int m = -f.half_width; m <= f.half_width; ++m) {
              const float w = interior[m + f.half_width];
              if (w == 0.f) continue;
              if (n_taps >= tdb::kMatMaxTaps) { linear = false; break; }
              a.tap_var[n_taps] = (short)f.var; a.tap_axis[n_taps] = (short)f.axis; a.tap_m[n_taps] = (short)m;
              a.tap_w[n_taps] = tm.coeff * w;
              ++n_taps;
            }
            if (f.axis == 0) ey = f.n_edge > ey ? f.n_edge : ey; else ex = f.n_edge > ex ? f.n_edge : ex;
          }
        } else {
          linear = false;
        }
      }
    }
    a.tap_begin[desc->n_eq] = n_taps;
    a.frc_begin[desc->n_eq] = n_frc;
    a.linear = (linear && n_lin <= tdb::kMatMaxTaps) ? 1 : 0;
    a.n_lin = n_lin;
    a.lin1 = (a.linear && desc->n_eq == 1 && desc->n_var == 1 && n_taps <= 16) ? 1 : 0;
    a.edge_y = ey; a.edge_x = ex;
  }
  a.tiles_x = (desc->n1 + tdb::kMatTX - 1) / tdb::kMatTX;
  a.tiles_y = (desc->n0 + tdb::kMatTY - 1) / tdb::kMatTY;
  p->n_ctas = a.tiles_x * a.tiles_y;
  const int uy = tdb::kMatTY + 4 * hy, ux = tdb::kMatTX + 4 * hx, ry = tdb::kMatTY + 2 * hy, rx = tdb::kMatTX + 2 * hx;
  p->smem = ((size_t)desc->n_var * uy * ux + (size_t)desc->n_fields * ry * rx) * sizeof(float);
  if (p->smem > 200 * 1024) { delete p; return mat_invalid("tile does not fit shared memory"); }
  MCU(cudaMalloc(&p->d_band, sizeof(float) * (n_band > 0 ? n_band : 1)));
  MCU(cudaMemcpy(p->d_band, band, sizeof(float) * n_band, cudaMemcpyHostToDevice));
  MCU(cudaMalloc(&p->d_terms, sizeof(tdb200_term) * (n_terms > 0 ? n_terms : 1)));
  if (n_terms) MCU(cudaMemcpy(p->d_terms, terms, sizeof(tdb200_term) * n_terms, cudaMemcpyHostToDevice));
  MCU(cudaMalloc(&p->d_factors, sizeof(tdb200_factor) * (n_factors > 0 ? n_factors : 1)));
  if (n_factors) MCU(cudaMemcpy(p->d_factors, factors, sizeof(tdb200_factor) * n_factors, cudaMemcpyHostToDevice));
  MCU(cudaMalloc(&p->d_part_loss, sizeof(double) * (size_t)p->n_ctas * desc->n_eq));
  a.band = p->d_band; a.terms = p->d_terms; a.factors = p->d_factors; a.part_loss = p->d_part_loss;
  tdb::MatBcArgs& b = p->bc;
  b.n_var = desc->n_var; b.n0 = desc->n0; b.n1 = desc->n1; b.n_fields = desc->n_fields; b.n_eq = desc->n_eq;
  b.band = p->d_band; b.terms = p->d_terms; b.factors = p->d_factors;
  if (p->smem > 48 * 1024)
    MCU(cudaFuncSetAttribute(tdb::mat_residual_adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
  *out = p;
  return TDB200_OK;
}

int tdb200_mat_plan_set_coeffs(tdb200_mat_plan* p, const float* coeffs_dev, int64_t n_coeffs) {
  if (!p) return mat_invalid("null plan");
  (void)n_coeffs;
  p->args.coeffs = coeffs_dev;
  p->bc.coeffs = coeffs_dev;
  return TDB200_OK;
}

int tdb200_mat_plan_set_bcs(tdb200_mat_plan* p, int32_t n_bcs, const tdb200_mat_bc* bcs, const int32_t* cells_dev,
                            const float* targets_dev, int32_t n_slots, const double* slot_lambda,
                            const double* slot_len) {
  if (!p || !slot_lambda || !slot_len || n_bcs < 0) return mat_invalid("null argument");
  const int n_eq = p->desc.n_eq;
  if (n_slots < n_eq || n_slots > 32) return mat_invalid("n_slots out of range");
  MCU(cudaSetDevice(p->device));
  p->n_slots = n_slots;
  p->n_bc_slots = n_slots - n_eq;
  std::vector<long long> begin(n_bcs + 1, 0);
  for (int i = 0; i < n_bcs; ++i) {
    if (bcs[i].K < 1 || bcs[i].K > 4 || bcs[i].slot < 0 || bcs[i].slot >= p->n_bc_slots) return mat_invalid("bad boundary descriptor");
    begin[i + 1] = begin[i] + bcs[i].n_rows;
  }
  p->n_bc_rows = begin[n_bcs];
  cudaFree(p->d_bcs); cudaFree(p->d_bc_row_begin); cudaFree(p->d_slot_scale); cudaFree(p->d_slot_lambda);
  cudaFree(p->d_slot_len); cudaFree(p->d_bc_sum);
  MCU(cudaMalloc(&p->d_bcs, sizeof(tdb200_mat_bc) * (n_bcs > 0 ? n_bcs : 1)));
  if (n_bcs) MCU(cudaMemcpy(p->d_bcs, bcs, sizeof(tdb200_mat_bc) * n_bcs, cudaMemcpyHostToDevice));
  MCU(cudaMalloc(&p->d_bc_row_begin, sizeof(long long) * (n_bcs + 1)));
  MCU(cudaMemcpy(p->d_bc_row_begin, begin.data(), sizeof(long long) * (n_bcs + 1), cudaMemcpyHostToDevice));
  std::vector<float> scale(p->n_bc_slots > 0 ? p->n_bc_slots : 1, 0.f);
  for (int s = 0; s < p->n_bc_slots; ++s) scale[s] = (float)(slot_lambda[n_eq + s] / slot_len[n_eq + s]);
  MCU(cudaMalloc(&p->d_slot_scale, sizeof(float) * scale.size()));
  MCU(cudaMemcpy(p->d_slot_scale, scale.data(), sizeof(float) * scale.size(), cudaMemcpyHostToDevice));
  MCU(cudaMalloc(&p->d_slot_lambda, sizeof(double) * n_slots));
  MCU(cudaMemcpy(p->d_slot_lambda, slot_lambda, sizeof(double) * n_slots, cudaMemcpyHostToDevice));
  MCU(cudaMalloc(&p->d_slot_len, sizeof(double) * n_slots));
  MCU(cudaMemcpy(p->d_slot_len, slot_len, sizeof(double) * n_slots, cudaMemcpyHostToDevice));
  MCU(cudaMalloc(&p->d_bc_sum, sizeof(double) * (p->n_bc_slots > 0 ? p->n_bc_slots : 1)));
  for (int e = 0; e < n_eq; ++e) p->args.eq_scale[e] = (float)(slot_lambda[e] / slot_len[e]);
  tdb::MatBcArgs& b = p->bc;
  b.bcs = p->d_bcs; b.n_bcs = n_bcs; b.bc_row_begin = p->d_bc_row_begin; b.cells = cells_dev; b.targets = targets_dev;
  b.slot_scale = p->d_slot_scale; b.slot_sum = p->d_bc_sum;
  p->bcs_set = true;
  return TDB200_OK;
}

static int mat_run(tdb200_mat_plan* p, const float* u, float* grad, float* op_out, float* bval_out, float* out,
                   void* stream) {
  if (!p || !u || !out) return mat_invalid("null argument");
  if (!p->bcs_set) return mat_invalid("tdb200_mat_plan_set_bcs was not called");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MCU(cudaSetDevice(p->device));
  tdb::MatArgs a = p->args;
  a.u = u; a.grad = grad; a.op_out = op_out;
  MCU(cudaMemsetAsync(p->d_bc_sum, 0, sizeof(double) * (p->n_bc_slots > 0 ? p->n_bc_slots : 1), s));
  dim3 grid(a.tiles_x, a.tiles_y);
  tdb::mat_residual_adjoint_kernel<<<grid, tdb::kMatThreads, p->smem, s>>>(a);
  MCU(cudaGetLastError());
  if (p->n_bc_rows > 0) {
    tdb::MatBcArgs b = p->bc;
    b.u = u; b.grad = grad; b.bval_out = bval_out;
    const int blocks = (int)((p->n_bc_rows + 127) / 128);
    tdb::mat_bc_kernel<<<blocks < 1184 ? blocks : 1184, 128, 0, s>>>(b);
    MCU(cudaGetLastError());
  }
  tdb::mat_finalize_kernel<<<1, 256, 0, s>>>(p->d_part_loss, p->n_ctas, p->desc.n_eq,
                                            (double)p->desc.n0 * (double)p->desc.n1, p->d_bc_sum, p->n_bc_slots,
                                            p->d_slot_lambda, p->d_slot_len, out);
  MCU(cudaGetLastError());
  return TDB200_OK;
}

int tdb200_mat_loss_grad(tdb200_mat_plan* p, const float* u_dev, float* grad_dev, float* out_dev, void* stream) {
  if (!grad_dev) return mat_invalid("null gradient buffer");
  return mat_run(p, u_dev, grad_dev, nullptr, nullptr, out_dev, stream);
}

int tdb200_mat_eval_fields(tdb200_mat_plan* p, const float* u_dev, float* op_dev, float* bval_dev, float* out_dev,
                           void* stream) {
  return mat_run(p, u_dev, nullptr, op_dev, bval_dev, out_dev, stream);
}

int64_t tdb200_mat_plan_out_size(const tdb200_mat_plan* p) { return p ? 2 + p->n_slots : 0; }
int32_t tdb200_mat_plan_launches_per_call(const tdb200_mat_plan* p) { return p ? (p->n_bc_rows > 0 ? 3 : 2) : 0; }

void tdb200_mat_plan_destroy(tdb200_mat_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaFree(p->d_band); cudaFree(p->d_terms); cudaFree(p->d_factors); cudaFree(p->d_bcs); cudaFree(p->d_bc_row_begin);
  cudaFree(p->d_slot_scale); cudaFree(p->d_slot_lambda); cudaFree(p->d_slot_len); cudaFree(p->d_part_loss);
  cudaFree(p->d_bc_sum);
  delete p;
}

}  // extern "C"
