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
#include <cuda.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "peer_dev.cuh"

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
  int lin1;                                // 1: single equation, single field, <= 16 taps, reach <= 4 -> mat_lin1_kernel
  float l1_fconst;                         // lin1: sum of the constant forcing terms
  const float* l1_fbuf[2];                 // lin1: up to two forcing buffers (absolute pointers, set per call)
  float cx_wy[9], cx_wx[9], cx_wc;         // cross kernel: weights by offset (index offset + reach), merged centre
  unsigned int* tile_ctr;                  // persistent kernel: dynamic tile counter (zero on entry)
  int row_lo, row_hi;                      // rows whose residuals enter the loss (slab decomposition: the owned rows of
                                           // an extended slab; seeds are still formed on the halo rows around them)
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

// ---- register-tap kernel: one linear constant-coefficient equation on one field -----------------------
// (Poisson, heat, wave ... in mat mode.)  128 x 32 tiles, compile-time shared-memory pitches so that every tap of
// every unrolled cell is one LDS with an immediate offset; interior CTAs run without any per-cell bounds or
// edge-row checks.  Cells whose stencil rows are special (one-sided rows near the domain edge) evaluate the
// banded operators directly.  HBM traffic per cell: read u, read the forcing buffer(s), write the gradient.
constexpr int kL1TY = 32, kL1TX = 128, kL1MaxH = 4;         // tile; composite stencil reach per axis <= 4
constexpr int kL1PU = kL1TX + 4 * kL1MaxH;                  // pitch of the u region (tile + 2 halos)
constexpr int kL1PR = kL1TX + 2 * kL1MaxH;                  // pitch of the seed region (tile + halo)
constexpr int kL1UY = kL1TY + 4 * kL1MaxH, kL1RY = kL1TY + 2 * kL1MaxH;
constexpr size_t kL1Smem = (size_t)(kL1UY * kL1PU + kL1RY * kL1PR) * sizeof(float);

template <int NT, bool INTERIOR>
__device__ __forceinline__ float mat_lin1_tile(const MatArgs& a, float* __restrict__ us, float* __restrict__ ss,
                                               int ty0, int tx0) {
  const int hy = a.hy, hx = a.hx;
  const int uy = kL1TY + 4 * hy, ux = kL1TX + 4 * hx, ry = kL1TY + 2 * hy, rx = kL1TX + 2 * hx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- phase 1: u tile + 2 halos -> shared memory (zero outside the domain) ----
#pragma unroll
  for (int i = 0; i < kL1UY / 8; ++i) {
    const int ly = warp + 8 * i;
    if (ly < uy) {
      const int gy = ty0 - 2 * hy + ly;
      const bool rowok = INTERIOR || (gy >= 0 && gy < a.n0);
      const long long rbase = (long long)gy * a.n1 + (tx0 - 2 * hx);
#pragma unroll
      for (int j = 0; j < (kL1PU + 31) / 32; ++j) {
        const int lx = lane + 32 * j;
        if (lx < ux) {
          const int gx = tx0 - 2 * hx + lx;
          float v = 0.f;
          if (INTERIOR || (rowok && gx >= 0 && gx < a.n1)) v = __ldg(a.u + rbase + lx);
          us[ly * kL1PU + lx] = v;
        }
      }
    }
  }
  float tw[NT];
  int tou[NT], tos[NT];
  const int nt = a.tap_begin[1];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const bool on = t < nt;
    tw[t] = on ? a.tap_w[t] : 0.f;
    const int m = on ? a.tap_m[t] : 0;
    const bool ax0 = on && a.tap_axis[t] == 0;
    tou[t] = ax0 ? m * kL1PU : m;
    tos[t] = ax0 ? -m * kL1PR : -m;
  }
  const float scale2 = 2.f * a.eq_scale[0];
  const float fc0 = a.l1_fconst;
  const float* __restrict__ f0 = a.l1_fbuf[0];
  const float* __restrict__ f1 = a.l1_fbuf[1];
  const int zy = a.edge_y, zx = a.edge_x;                  // rows / columns with their own coefficients
  float lacc = 0.f;
  __syncthreads();
  // ---- phase 2: residual seeds on tile + halo ----
#pragma unroll
  for (int i = 0; i < kL1RY / 8; ++i) {
    const int ly = warp + 8 * i;
    if (ly < ry) {
      const int gy = ty0 - hy + ly;
      const long long rbase = (long long)gy * a.n1 + (tx0 - hx);
#pragma unroll
      for (int j = 0; j < (kL1PR + 31) / 32; ++j) {
        const int lx = lane + 32 * j;
        if (lx < rx) {
          const int gx = tx0 - hx + lx;
          float seed = 0.f;
          if (INTERIOR || (gy >= 0 && gy < a.n0 && gx >= 0 && gx < a.n1)) {
            float res = fc0;
            if (f0) res += __ldg(f0 + rbase + lx);
            if (f1) res += __ldg(f1 + rbase + lx);
            const float* uc = us + (ly + hy) * kL1PU + lx + hx;
            if (INTERIOR || (gy >= zy && gy < a.n0 - zy && gx >= zx && gx < a.n1 - zx)) {
#pragma unroll
              for (int t = 0; t < NT; ++t) res = fmaf(tw[t], uc[tou[t]], res);
            } else {
              for (int t = 0; t < a.n_lin; ++t)
                res = fmaf(a.lin_c[t], field_value(a, a.fld[a.lin_q[t]], us, kL1PU, 0, ly + hy, lx + hx, gy, gx), res);
            }
            if (ly >= hy && ly < hy + kL1TY && lx >= hx && lx < hx + kL1TX && gy >= a.row_lo && gy < a.row_hi) lacc = fmaf(res, res, lacc);
            seed = scale2 * res;
          }
          ss[ly * kL1PR + lx] = seed;
        }
      }
    }
  }
  __syncthreads();
  if (!a.grad) return lacc;
  // ---- phase 3: transposed stencil -> gradient of the tile ----
  const int zy3 = zy + hy, zx3 = zx + hx;                  // cells that gather from a special row
#pragma unroll
  for (int i = 0; i < kL1TY / 8; ++i) {
    const int cy = warp + 8 * i, gy = ty0 + cy;
#pragma unroll
    for (int j = 0; j < kL1TX / 32; ++j) {
      const int cx = lane + 32 * j, gx = tx0 + cx;
      if (INTERIOR || (gy < a.n0 && gx < a.n1)) {
        const float* sc = ss + (cy + hy) * kL1PR + cx + hx;
        float g = 0.f;
        if (INTERIOR || (gy >= zy3 && gy < a.n0 - zy3 && gx >= zx3 && gx < a.n1 - zx3)) {
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
                if (yy >= 0 && yy < a.n0) s = fmaf(band_coef(a.band, f, a.n0, yy, -m), sc[m * kL1PR], s);
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
  }
  return lacc;
}

template <int NT>
__global__ void __launch_bounds__(256, 3) mat_lin1_kernel(const MatArgs a) {
  extern __shared__ __align__(16) float sm_l1[];
  float* us = sm_l1;
  float* ss = sm_l1 + kL1UY * kL1PU;
  __shared__ double red[8];
  const int ty0 = blockIdx.y * kL1TY, tx0 = blockIdx.x * kL1TX;
  const bool interior = ty0 - 2 * a.hy >= a.edge_y && ty0 + kL1TY + 2 * a.hy <= a.n0 - a.edge_y &&
                        tx0 - 2 * a.hx >= a.edge_x && tx0 + kL1TX + 2 * a.hx <= a.n1 - a.edge_x;
  const float lacc = interior ? mat_lin1_tile<NT, true>(a, us, ss, ty0, tx0) : mat_lin1_tile<NT, false>(a, us, ss, ty0, tx0);
  double v = (double)lacc;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    a.part_loss[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

template <int NT>
static cudaError_t launch_mat_lin1_nt(const MatArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mat_lin1_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kL1Smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((a.n1 + kL1TX - 1) / kL1TX, (a.n0 + kL1TY - 1) / kL1TY);
  mat_lin1_kernel<NT><<<grid, 256, kL1Smem, s>>>(a);
  return cudaGetLastError();
}
static cudaError_t launch_mat_lin1(const MatArgs& a, cudaStream_t s) {
  const int nt = a.tap_begin[1];
  if (nt <= 4) return launch_mat_lin1_nt<4>(a, s);
  if (nt <= 6) return launch_mat_lin1_nt<6>(a, s);
  if (nt <= 8) return launch_mat_lin1_nt<8>(a, s);
  if (nt <= 12) return launch_mat_lin1_nt<12>(a, s);
  return launch_mat_lin1_nt<16>(a, s);
}
static int mat_lin1_ctas(const MatArgs& a) {
  return ((a.n1 + kL1TX - 1) / kL1TX) * ((a.n0 + kL1TY - 1) / kL1TY);
}

// ---- vectorised cross-stencil kernel -----------------------------------------------------------------------
// Same contract as mat_lin1_kernel for operators whose composite interior stencil is a cross with reach HY / HX
// (compile-time masks of the non-zero offsets) on grids with n1 % 4 == 0: the u tile is staged with 16-byte
// cp.async (zero fill outside the domain), every thread works on float4 column groups (LDS.128 / LDG.128 /
// STG.128), the forcing values of the residual pass are fetched into registers before the staging wait so that
// both HBM streams are in flight together.  Cells whose stencil rows are special (one-sided rows near the domain
// edge) are recomputed afterwards by compact fix-up passes with the banded operators (boundary CTAs only).
constexpr int kCxTY = 32, kCxTX = 128;
constexpr int kCxPU = kCxTX + 16;                           // u region columns [tx0 - 8, tx0 + 136)
constexpr int kCxPR = kCxTX + 8;                            // seed region columns [tx0 - 4, tx0 + 132)
constexpr int kCxQU = kCxPU / 4, kCxQR = kCxPR / 4;         // float4 groups per row

__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(src_bytes) : "memory");
}

// sum over the cross taps for the 4 cells of one column group.  REV: transposed stencil, g[c] += w[d] * s[c - d].
template <int HY, int HX, unsigned MY, unsigned MX, bool REV, int PITCH>
__device__ __forceinline__ void cx_apply(const float* __restrict__ c, const float* wy, const float* wx, float wc, float* r) {
  const float4 c0 = *reinterpret_cast<const float4*>(c);
  float w[12];
  w[4] = c0.x; w[5] = c0.y; w[6] = c0.z; w[7] = c0.w;
  if (MX != 0) {
    const float4 cl = *reinterpret_cast<const float4*>(c - 4), cr = *reinterpret_cast<const float4*>(c + 4);
    w[0] = cl.x; w[1] = cl.y; w[2] = cl.z; w[3] = cl.w;
    w[8] = cr.x; w[9] = cr.y; w[10] = cr.z; w[11] = cr.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fmaf(wc, w[4 + i], r[i]);
#pragma unroll
  for (int dx = -HX; dx <= HX; ++dx) {
    if (dx == 0 || !(MX & (1u << (dx + HX)))) continue;
    const float wgt = wx[dx + HX];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fmaf(wgt, w[4 + i + (REV ? -dx : dx)], r[i]);
  }
#pragma unroll
  for (int dy = -HY; dy <= HY; ++dy) {
    if (dy == 0 || !(MY & (1u << (dy + HY)))) continue;
    const float wgt = wy[dy + HY];
    const float4 v = *reinterpret_cast<const float4*>(c + (REV ? -dy : dy) * PITCH);
    r[0] = fmaf(wgt, v.x, r[0]); r[1] = fmaf(wgt, v.y, r[1]); r[2] = fmaf(wgt, v.z, r[2]); r[3] = fmaf(wgt, v.w, r[3]);
  }
}

// in-domain cells of the seed region whose stencil rows are special: (a) whole special rows, warp-uniform;
// (b) special columns of the remaining rows, compact enumeration
template <int HY, int HX>
__device__ __forceinline__ void cx_fix_seeds(const MatArgs& a, const float* __restrict__ us, float* __restrict__ ss,
                                             int ty0, int tx0, float& lacc) {
  constexpr int RY = kCxTY + 2 * HY;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = a.n0, n1 = a.n1, zy = a.edge_y, zx = a.edge_x;
  const float* __restrict__ f0 = a.l1_fbuf[0];
  const float* __restrict__ f1 = a.l1_fbuf[1];
  const float fc0 = a.l1_fconst, scale2 = 2.f * a.eq_scale[0];
  auto special_seed = [&](int ly, int lx) {                // (ly, lx): seed-region coordinates
    const int gy = ty0 - HY + ly, gx = tx0 - 4 + lx;
    const size_t cell = (size_t)gy * n1 + gx;
    float res = fc0;
    if (f0) res += __ldg(f0 + cell);
    if (f1) res += __ldg(f1 + cell);
    for (int t = 0; t < a.n_lin; ++t)
      res = fmaf(a.lin_c[t], field_value(a, a.fld[a.lin_q[t]], us, kCxPU, 0, ly + HY, lx + 4, gy, gx), res);
    if (ly >= HY && ly < HY + kCxTY && lx >= 4 && lx < 4 + kCxTX && gy >= a.row_lo && gy < a.row_hi) lacc = fmaf(res, res, lacc);
    ss[ly * kCxPR + lx] = scale2 * res;
  };
  for (int ly = warp; ly < RY; ly += (int)(blockDim.x >> 5)) {
    const int gy = ty0 - HY + ly;
    if (gy < 0 || gy >= n0 || (gy >= zy && gy < n0 - zy)) continue;
    for (int lx = 4 - HX + lane; lx < 4 + kCxTX + HX; lx += 32) {
      const int gx = tx0 - 4 + lx;
      if (gx >= 0 && gx < n1) special_seed(ly, lx);
    }
  }
  for (int k = tid; k < RY * 2 * zx; k += (int)blockDim.x) {
    const int ly = k / (2 * zx), ci = k - ly * (2 * zx);
    const int gy = ty0 - HY + ly, gx = ci < zx ? ci : n1 - 2 * zx + ci;
    const int lx = gx - (tx0 - 4);
    if (gy >= zy && gy < n0 - zy && lx >= 4 - HX && lx < 4 + kCxTX + HX) special_seed(ly, lx);
  }
}

// cells of the tile whose transposed stencil gathers from a special row
template <int HY, int HX>
__device__ __forceinline__ void cx_fix_grad(const MatArgs& a, const float* __restrict__ ss, int ty0, int tx0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = a.n0, n1 = a.n1;
  const int zy3 = a.edge_y + HY, zx3 = a.edge_x + HX;
  auto special_grad = [&](int cy, int cx) {
    const int gy = ty0 + cy, gx = tx0 + cx;
    const float* sc = ss + (cy + HY) * kCxPR + cx + 4;
    float g = 0.f;
    for (int t = 0; t < a.n_lin; ++t) {
      const tdb200_mat_field& f = a.fld[a.lin_q[t]];
      float s = 0.f;
      if (f.order == 0) s = sc[0];
      else if (f.axis == 0) {
        for (int m = -f.half_width; m <= f.half_width; ++m) {
          const int yy = gy + m;
          if (yy >= 0 && yy < n0) s = fmaf(band_coef(a.band, f, n0, yy, -m), sc[m * kCxPR], s);
        }
      } else {
        for (int m = -f.half_width; m <= f.half_width; ++m) {
          const int xx = gx + m;
          if (xx >= 0 && xx < n1) s = fmaf(band_coef(a.band, f, n1, xx, -m), sc[m], s);
        }
      }
      g = fmaf(a.lin_c[t], s, g);
    }
    a.grad[(size_t)gy * n1 + gx] = g;
  };
  for (int cy = warp; cy < kCxTY; cy += (int)(blockDim.x >> 5)) {
    const int gy = ty0 + cy;
    if (gy >= n0 || (gy >= zy3 && gy < n0 - zy3)) continue;
    for (int cx = lane; cx < kCxTX; cx += 32)
      if (tx0 + cx < n1) special_grad(cy, cx);
  }
  for (int k = tid; k < kCxTY * 2 * zx3; k += (int)blockDim.x) {
    const int cy = k / (2 * zx3), ci = k - cy * (2 * zx3);
    const int gy = ty0 + cy, gx = ci < zx3 ? ci : n1 - 2 * zx3 + ci;
    const int cx = gx - tx0;
    if (gy < n0 && gy >= zy3 && gy < n0 - zy3 && cx >= 0 && cx < kCxTX) special_grad(cy, cx);
  }
}

template <int HY, int HX, unsigned MY, unsigned MX, bool INTERIOR>
__device__ __forceinline__ float mat_cross_tile(const MatArgs& a, float* __restrict__ us, float* __restrict__ ss,
                                                int ty0, int tx0) {
  constexpr int UY = kCxTY + 4 * HY, RY = kCxTY + 2 * HY;
  constexpr int N2 = (RY * kCxQR + 255) / 256;               // residual-pass items per thread
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = a.n0, n1 = a.n1;
  // ---- phase 1: stage u (tile + 2 halos), asynchronously ----
  for (int idx = tid; idx < UY * kCxQU; idx += 256) {
    const int ly = idx / kCxQU, q = idx - ly * kCxQU;
    const int gy = ty0 - 2 * HY + ly, gx = tx0 - 8 + 4 * q;
    const bool ok = INTERIOR || (gy >= 0 && gy < n0 && gx >= 0 && gx < n1);
    cp_async16(us + ly * kCxPU + 4 * q, ok ? a.u + (size_t)gy * n1 + gx : a.u, ok ? 16 : 0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // ---- forcing values of this thread's residual items -> registers (overlaps the staging) ----
  const float* __restrict__ f0 = a.l1_fbuf[0];
  const float* __restrict__ f1 = a.l1_fbuf[1];
  float4 fv[N2];
#pragma unroll
  for (int k = 0; k < N2; ++k) {
    const int idx = tid + 256 * k;
    const int ly = idx / kCxQR, q = idx - ly * kCxQR;
    const int gy = ty0 - HY + ly, gx = tx0 - 4 + 4 * q;
    fv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < RY * kCxQR && (INTERIOR || (gy >= 0 && gy < n0 && gx >= 0 && gx < n1))) {
      const size_t cell = (size_t)gy * n1 + gx;
      if (f0) fv[k] = __ldg(reinterpret_cast<const float4*>(f0 + cell));
      if (f1) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(f1 + cell));
        fv[k].x += t.x; fv[k].y += t.y; fv[k].z += t.z; fv[k].w += t.w;
      }
    }
  }
  float wy[2 * HY + 1], wx[2 * HX + 1];
#pragma unroll
  for (int i = 0; i <= 2 * HY; ++i) wy[i] = a.cx_wy[i];
#pragma unroll
  for (int i = 0; i <= 2 * HX; ++i) wx[i] = a.cx_wx[i];
  const float wc = a.cx_wc, fc0 = a.l1_fconst, scale2 = 2.f * a.eq_scale[0];
  const int zy = a.edge_y, zx = a.edge_x;
  float lacc = 0.f;
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // ---- phase 2: residual seeds on tile + halo ----
#pragma unroll
  for (int k = 0; k < N2; ++k) {
    const int idx = tid + 256 * k;
    if (idx < RY * kCxQR) {
      const int ly = idx / kCxQR, q = idx - ly * kCxQR;
      float r[4] = {fv[k].x + fc0, fv[k].y + fc0, fv[k].z + fc0, fv[k].w + fc0};
      cx_apply<HY, HX, MY, MX, false, kCxPU>(us + (ly + HY) * kCxPU + 4 * q + 4, wy, wx, wc, r);
      const bool core = ly >= HY && ly < HY + kCxTY && q >= 1 && q <= kCxTX / 4 && ty0 - HY + ly >= a.row_lo &&
                        ty0 - HY + ly < a.row_hi;
      float sd[4];
      if (INTERIOR) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { sd[i] = scale2 * r[i]; if (core) lacc = fmaf(r[i], r[i], lacc); }
      } else {
        const int gy = ty0 - HY + ly, gx = tx0 - 4 + 4 * q;
        const bool rowreg = gy >= zy && gy < n0 - zy;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool reg = rowreg && gx + i >= zx && gx + i < n1 - zx;   // regular interior row of the operators
          sd[i] = reg ? scale2 * r[i] : 0.f;
          if (core && reg) lacc = fmaf(r[i], r[i], lacc);
        }
      }
      *reinterpret_cast<float4*>(ss + ly * kCxPR + 4 * q) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    }
  }
  if (!INTERIOR) {
    __syncthreads();
    cx_fix_seeds<HY, HX>(a, us, ss, ty0, tx0, lacc);
  }
  __syncthreads();
  if (!a.grad) return lacc;
  // ---- phase 3: transposed stencil -> gradient of the tile ----
#pragma unroll
  for (int i = 0; i < kCxTY / 8; ++i) {
    const int cy = warp + 8 * i, gy = ty0 + cy, gx = tx0 + 4 * lane;
    if (INTERIOR || (gy < n0 && gx < n1)) {
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      cx_apply<HY, HX, MY, MX, true, kCxPR>(ss + (cy + HY) * kCxPR + 4 * lane + 4, wy, wx, wc, g);
      *reinterpret_cast<float4*>(a.grad + (size_t)gy * n1 + gx) = make_float4(g[0], g[1], g[2], g[3]);
    }
  }
  if (!INTERIOR) {
    __syncthreads();                                         // the fix-up overwrites cells stored above
    cx_fix_grad<HY, HX>(a, ss, ty0, tx0);
  }
  return lacc;
}

template <int HY, int HX, unsigned MY, unsigned MX>
__global__ void __launch_bounds__(256, 3) mat_cross_kernel(const MatArgs a) {
  constexpr int UY = kCxTY + 4 * HY;
  extern __shared__ __align__(16) float sm_cx[];
  float* us = sm_cx;
  float* ss = sm_cx + UY * kCxPU;
  __shared__ double red[8];
  const int ty0 = blockIdx.y * kCxTY, tx0 = blockIdx.x * kCxTX;
  const bool interior = ty0 - 2 * HY >= a.edge_y && ty0 + kCxTY + 2 * HY <= a.n0 - a.edge_y &&
                        tx0 - 8 >= a.edge_x && tx0 + kCxTX + 8 <= a.n1 - a.edge_x;
  const float lacc = interior ? mat_cross_tile<HY, HX, MY, MX, true>(a, us, ss, ty0, tx0)
                              : mat_cross_tile<HY, HX, MY, MX, false>(a, us, ss, ty0, tx0);
  double v = (double)lacc;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    a.part_loss[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

template <int HY, int HX, unsigned MY, unsigned MX>
static cudaError_t launch_mat_cross_t(const MatArgs& a, cudaStream_t s) {
  constexpr size_t smem = (size_t)((kCxTY + 4 * HY) * kCxPU + (kCxTY + 2 * HY) * kCxPR) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mat_cross_kernel<HY, HX, MY, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((a.n1 + kCxTX - 1) / kCxTX, (a.n0 + kCxTY - 1) / kCxTY);
  mat_cross_kernel<HY, HX, MY, MX><<<grid, 256, smem, s>>>(a);
  return cudaGetLastError();
}
// instantiated cross shapes: (reach y, reach x, mask y, mask x); mask bit (d + reach) <=> offset d is non-zero
#define TDB_CROSS_SHAPES(X) \
  X(2, 2, 0x11u, 0x11u) X(1, 2, 0x5u, 0x11u) X(2, 1, 0x11u, 0x5u) X(1, 1, 0x5u, 0x5u) X(4, 4, 0x1EFu, 0x1EFu) \
  X(2, 0, 0x11u, 0x0u) X(0, 2, 0x0u, 0x11u)
static bool mat_cross_supported(int hy, int hx, unsigned my, unsigned mx) {
#define X(A, B, C, D) if (hy == A && hx == B && (my & ~C) == 0 && (mx & ~D) == 0) return true;
  TDB_CROSS_SHAPES(X)
#undef X
  return false;
}
static cudaError_t launch_mat_cross(const MatArgs& a, int hy, int hx, unsigned my, unsigned mx, cudaStream_t s) {
#define X(A, B, C, D) if (hy == A && hx == B && (my & ~C) == 0 && (mx & ~D) == 0) return launch_mat_cross_t<A, B, C, D>(a, s);
  TDB_CROSS_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}
static int mat_cross_ctas(const MatArgs& a) {
  return ((a.n1 + kCxTX - 1) / kCxTX) * ((a.n0 + kCxTY - 1) / kCxTY);
}

// ---- persistent TMA cross-stencil kernel -----------------------------------------------------------------
// The production path for BASELINE config 4.  One CTA pair per SM walks over the 128 x 32 tiles; a single thread
// stages tile t+1 (the u box with both halos and the forcing box with one halo, out-of-range elements zero
// filled by the TMA unit) with two cp.async.bulk.tensor.2d loads completing on an mbarrier while all threads work
// on tile t: the HBM stream never stops, nothing is staged through registers and no thread computes a load address.
constexpr int kCtThreads = 512, kCtWarps = kCtThreads / 32;

__device__ __forceinline__ uint32_t ct_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ct_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ct_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ct_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ct_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = ct_smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void ct_tma_load_2d(float* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(ct_smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(ct_smem_u32(bar)) : "memory");
}

// per stage: u box [UY][PU] then forcing box [RY][PR]; both sub-buffers start 128-byte aligned (TMA destination)
__host__ __device__ constexpr size_t ct_round32(size_t n) { return (n + 31) / 32 * 32; }
template <int HY, int HX>
__host__ __device__ constexpr size_t ct_u_floats() { return ct_round32((size_t)(kCxTY + 4 * HY) * kCxPU); }
template <int HY, int HX>
__host__ __device__ constexpr size_t ct_stage_floats() { return ct_u_floats<HY, HX>() + ct_round32((size_t)(kCxTY + 2 * HY) * kCxPR); }
template <int HY, int HX>
constexpr size_t ct_smem_bytes() {
  return (2 * ct_stage_floats<HY, HX>() + ct_round32((size_t)(kCxTY + 2 * HY) * kCxPR)) * sizeof(float) + 128 /*align*/ + 64;
}

template <int HY, int HX, unsigned MY, unsigned MX>
__global__ void __launch_bounds__(kCtThreads, 2) mat_cross_tma_kernel(const MatArgs a, const __grid_constant__ CUtensorMap map_u,
                                                                      const __grid_constant__ CUtensorMap map_f) {
  constexpr int UY = kCxTY + 4 * HY, RY = kCxTY + 2 * HY;
  constexpr int N2 = (RY * kCxQR + kCtThreads - 1) / kCtThreads;
  constexpr uint32_t kTxBytes = (uint32_t)((UY * kCxPU + RY * kCxPR) * sizeof(float));
  extern __shared__ uint8_t sm_ct_raw[];
  const uint32_t s0_ = ct_smem_u32(sm_ct_raw);                 // align with pointer arithmetic on the __shared__ symbol:
  float* base = reinterpret_cast<float*>(sm_ct_raw + (((s0_ + 127u) & ~127u) - s0_));   // keeps LDS / STS addressing
  float* ss = base + 2 * ct_stage_floats<HY, HX>();          // seeds [RY][PR]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ss + ct_round32((size_t)RY * kCxPR));
  __shared__ double red[kCtThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = a.n0, n1 = a.n1;
  const int tiles_x = (n1 + kCxTX - 1) / kCxTX, n_tiles = tiles_x * ((n0 + kCxTY - 1) / kCxTY);
  const bool has_f = a.l1_fbuf[0] != nullptr;
  if (tid == 0) {
    ct_mbar_init(bars, 1);
    ct_mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto stage_tile = [&](int tile, int st) {                  // one thread
    float* us = base + st * ct_stage_floats<HY, HX>();
    float* fs = us + ct_u_floats<HY, HX>();
    const int ty0 = (tile / tiles_x) * kCxTY, tx0 = (tile % tiles_x) * kCxTX;
    ct_mbar_expect_tx(bars + st, has_f ? kTxBytes : (uint32_t)(UY * kCxPU * sizeof(float)));
    ct_tma_load_2d(us, &map_u, tx0 - 8, ty0 - 2 * HY, bars + st);
    if (has_f) ct_tma_load_2d(fs, &map_f, tx0 - 4, ty0 - HY, bars + st);
  };
  float wy[2 * HY + 1], wx[2 * HX + 1];
#pragma unroll
  for (int i = 0; i <= 2 * HY; ++i) wy[i] = a.cx_wy[i];
#pragma unroll
  for (int i = 0; i <= 2 * HX; ++i) wx[i] = a.cx_wx[i];
  const float wc = a.cx_wc, fc0 = a.l1_fconst, scale2 = 2.f * a.eq_scale[0];
  const int zy = a.edge_y, zx = a.edge_x;
  double dacc = 0.0;
  // this thread's items of the residual pass (seed-region row, float4 group): fixed for the whole kernel
  int off_u[N2], off_s[N2], lyq[N2];
  bool core[N2];
#pragma unroll
  for (int k = 0; k < N2; ++k) {
    const int idx = tid + kCtThreads * k;
    const int ly = idx / kCxQR, q = idx - ly * kCxQR;
    off_u[k] = idx < RY * kCxQR ? (ly + HY) * kCxPU + 4 * q + 4 : -1;
    off_s[k] = ly * kCxPR + 4 * q;
    lyq[k] = (ly << 8) | q;
    core[k] = ly >= HY && ly < HY + kCxTY && q >= 1 && q <= kCxTX / 4;
  }
  // dynamic tile scheduler: the first tile of a CTA is its block index, further tiles come from a global counter
  // (boundary tiles cost more than interior tiles); thread 0 publishes {tile, ty0, tx0, interior} of each stage
  __shared__ int4 tile_s[2];
  uint32_t phase[2] = {0u, 0u};
  auto publish = [&](int tile, int st) {                     // one thread
    const int ty0 = (tile / tiles_x) * kCxTY, tx0 = (tile % tiles_x) * kCxTX;
    const bool interior = ty0 - 2 * HY >= zy && ty0 + kCxTY + 2 * HY <= n0 - zy && tx0 - 8 >= zx && tx0 + kCxTX + 8 <= n1 - zx;
    tile_s[st] = make_int4(tile, ty0, tx0, interior ? 1 : 0);
    if (tile < n_tiles) stage_tile(tile, st);
  };
  if (tid == 0) publish(blockIdx.x, 0);
  __syncthreads();
  for (int st = 0;; st ^= 1) {
    const int4 ti = tile_s[st];
    if (ti.x >= n_tiles) break;
    if (tid == 0) publish((int)gridDim.x + (int)atomicAdd(a.tile_ctr, 1u), st ^ 1);   // read after this iteration's barriers
    const float* us = base + st * ct_stage_floats<HY, HX>();
    const float* fs = us + ct_u_floats<HY, HX>();
    const int ty0 = ti.y, tx0 = ti.z;
    const bool interior = ti.w != 0;
    const bool rows_in = ty0 >= a.row_lo && ty0 + kCxTY <= a.row_hi;      // tile entirely inside the loss window
    auto in_win = [&](int ly) { const int gy = ty0 - HY + ly; return gy >= a.row_lo && gy < a.row_hi; };
    ct_mbar_wait(bars + st, phase[st]);
    phase[st] ^= 1u;
    float lacc = 0.f;
    // ---- residual seeds on tile + halo ----
#pragma unroll
    for (int k = 0; k < N2; ++k) {
      if (off_u[k] >= 0) {
        float r[4] = {fc0, fc0, fc0, fc0};
        if (has_f) {
          const float4 fv = *reinterpret_cast<const float4*>(fs + off_s[k]);
          r[0] += fv.x; r[1] += fv.y; r[2] += fv.z; r[3] += fv.w;
        }
        cx_apply<HY, HX, MY, MX, false, kCxPU>(us + off_u[k], wy, wx, wc, r);
        float sd[4];
        if (interior) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { sd[i] = scale2 * r[i]; if (core[k] && (rows_in || in_win(lyq[k] >> 8))) lacc = fmaf(r[i], r[i], lacc); }
        } else {
          const int gy = ty0 - HY + (lyq[k] >> 8), gx = tx0 - 4 + 4 * (lyq[k] & 255);
          const bool rowreg = gy >= zy && gy < n0 - zy;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool reg = rowreg && gx + i >= zx && gx + i < n1 - zx;   // regular interior row of the operators
            sd[i] = reg ? scale2 * r[i] : 0.f;
            if (core[k] && reg && (rows_in || in_win(lyq[k] >> 8))) lacc = fmaf(r[i], r[i], lacc);
          }
        }
        *reinterpret_cast<float4*>(ss + off_s[k]) = make_float4(sd[0], sd[1], sd[2], sd[3]);
      }
    }
    if (!interior) {                                         // CTA-uniform
      __syncthreads();
      cx_fix_seeds<HY, HX>(a, us, ss, ty0, tx0, lacc);
    }
    dacc += (double)lacc;
    __syncthreads();
    // ---- transposed stencil -> gradient of the tile ----
    if (a.grad) {
#pragma unroll
      for (int i = 0; i < kCxTY / kCtWarps; ++i) {
        const int cy = warp + kCtWarps * i, gy = ty0 + cy, gx = tx0 + 4 * lane;
        if (gy < n0 && gx < n1) {
          float g[4] = {0.f, 0.f, 0.f, 0.f};
          cx_apply<HY, HX, MY, MX, true, kCxPR>(ss + (cy + HY) * kCxPR + 4 * lane + 4, wy, wx, wc, g);
          *reinterpret_cast<float4*>(a.grad + (size_t)gy * n1 + gx) = make_float4(g[0], g[1], g[2], g[3]);
        }
      }
      if (!interior) {
        __syncthreads();                                     // the fix-up overwrites cells stored above
        cx_fix_grad<HY, HX>(a, ss, ty0, tx0);
      }
    }
    __syncthreads();                                         // seeds and this stage's buffers are free again
  }
  for (int o = 16; o; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
  if (lane == 0) red[warp] = dacc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kCtThreads / 32; ++w) s += red[w];
    a.part_loss[blockIdx.x] = s;
  }
}

template <int HY, int HX, unsigned MY, unsigned MX>
static cudaError_t launch_mat_cross_tma_t(const MatArgs& a, const CUtensorMap& mu, const CUtensorMap& mf, int n_sms, int* n_ctas,
                                          cudaStream_t s) {
  constexpr size_t smem = ct_smem_bytes<HY, HX>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mat_cross_tma_kernel<HY, HX, MY, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int per_sm = smem * 2 <= 227 * 1024 ? 2 : 1;
  const int n_tiles = ((a.n1 + kCxTX - 1) / kCxTX) * ((a.n0 + kCxTY - 1) / kCxTY);
  int grid = n_sms * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  *n_ctas = grid;
  mat_cross_tma_kernel<HY, HX, MY, MX><<<grid, kCtThreads, smem, s>>>(a, mu, mf);
  return cudaGetLastError();
}
static cudaError_t launch_mat_cross_tma(const MatArgs& a, int hy, int hx, unsigned my, unsigned mx, const CUtensorMap& mu,
                                        const CUtensorMap& mf, int n_sms, int* n_ctas, cudaStream_t s) {
#define X(A, B, C, D) if (hy == A && hx == B && (my & ~C) == 0 && (mx & ~D) == 0) return launch_mat_cross_tma_t<A, B, C, D>(a, mu, mf, n_sms, n_ctas, s);
  TDB_CROSS_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}
// 2-D fp32 tensor map over a row-major [n0][n1] array with a [box_y][box_x] box, no swizzle, zero fill out of range
typedef CUresult (*tdb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_map_2d(CUtensorMap* map, const float* ptr, int n0, int n1, int box_y, int box_x) {
  static tdb_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tdb_encode_tiled_fn>(f);
  }
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)n1, (cuuint64_t)n0};
  const cuuint64_t gstr[1] = {(cuuint64_t)n1 * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_x, (cuuint32_t)box_y};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- register-marching cross-stencil kernel ---------------------------------------------------------------
// BASELINE config 4 (Poisson 4096 x 4096), second generation.  No shared-memory tiles, no block barriers, no 2-D halo:
// every WARP owns a column strip (32 lanes x float4 = 128 loaded columns, the inner 30 lanes = 120 columns are its
// output) and marches down a chunk of rows.  The last 2 HY + 1 + P rows of u, the last 2 HY + 1 rows of residual seeds
// and the forcing rows in flight live in REGISTERS (rings indexed at compile time: the row loop is unrolled by the
// ring length); x-neighbours come from the two adjacent lanes by warp shuffles.  Per row and thread: two 16-byte
// global loads (u, f) issued P rows ahead of their use (the HBM stream is kept full by register prefetch rather than
// by occupancy), <= 8 shuffles, ~45 FMAs, one 16-byte store.  Seeds of cells whose stencil rows are special (one-sided
// rows near the domain edge) are zero here; `mat_march_edge_kernel` adds their loss and gradient contributions.
constexpr int kMwWarps = 8, kMwThreads = kMwWarps * 32;
constexpr int kMwOutLanes = 30, kMwOutW = 4 * kMwOutLanes;   // lanes 1..30 own output columns: the composite stencil reaches
// 2 HX <= 4 columns = ONE float4 lane to each side (lanes 0 / 31 hold u and the two seed columns next to the strip)
constexpr unsigned kFullMask = 0xffffffffu;

// r[i] += sum_dx wx[dx] * row[x_i + dx] (REV: row[x_i - dx]) for the 4 cells of a thread; neighbours by shuffle
template <int H, unsigned M, bool REV>
__device__ __forceinline__ void mw_xtaps(const float4 c, const float* wx, float* r) {
  if (M == 0) return;
  float w[12];
  w[4] = c.x; w[5] = c.y; w[6] = c.z; w[7] = c.w;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bool need_l = false, need_r = false;
#pragma unroll
    for (int dx = -H; dx <= H; ++dx) {
      if (dx == 0 || !(M & (1u << (dx + H)))) continue;
      const int d = REV ? -dx : dx;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (4 + i + d == e) need_l = true;
        if (4 + i + d == 8 + e) need_r = true;
      }
    }
    w[e] = need_l ? __shfl_up_sync(kFullMask, w[4 + e], 1) : 0.f;
    w[8 + e] = need_r ? __shfl_down_sync(kFullMask, w[4 + e], 1) : 0.f;
  }
#pragma unroll
  for (int dx = -H; dx <= H; ++dx) {
    if (dx == 0 || !(M & (1u << (dx + H)))) continue;
    const float wgt = wx[dx + H];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fmaf(wgt, w[4 + i + (REV ? -dx : dx)], r[i]);
  }
}

// Cells near the domain edge.  The FRAME = cells within edge + reach of the domain edge, enumerated band by band (top
// and bottom rows at full width, then the left / right columns of the rows between).  Phase A (the leading CTAs of the
// mat_march_kernel launch): residual of every frame cell whose stencil rows are special (not "regular"), evaluated from
// u with the banded operators -> loss partial + compact seed buffer.  Phase B (mat_march_edge_kernel, after the march
// launch, before the boundary rows): every frame cell gathers coef(row of the special neighbour, this cell) * seed over
// its special in-domain neighbours and adds it to the gradient the march kernel wrote (regular rows only).  Each
// gradient cell has one owner: bit-reproducible.  All fields of a march-kernel plan have half_width <= 2.
constexpr int kMwMaxHw = 2;
struct MwFrame {
  int n0, n1, zy, zx, top, bot, mid, left, right;
  int n_band, total;                                         // < 2^31: the frame is a few rows / columns wide
  __device__ __forceinline__ MwFrame(const MatArgs& a, int zy3, int zx3) {
    n0 = a.n0; n1 = a.n1; zy = a.edge_y; zx = a.edge_x;
    top = min(zy3, n0); bot = min(zy3, n0 - top); mid = n0 - top - bot;
    left = min(zx3, n1); right = min(zx3, n1 - left);
    n_band = (top + bot) * n1; total = n_band + mid * (left + right);
  }
  __device__ __forceinline__ bool regular(int gy, int gx) const { return gy >= zy && gy < n0 - zy && gx >= zx && gx < n1 - zx; }
  __device__ __forceinline__ void cell(int idx, int& gy, int& gx) const {
    if (idx < n_band) {
      const int i = idx / n1;
      gx = idx - i * n1;
      gy = i < top ? i : n0 - bot + (i - top);
    } else {
      const int k = idx - n_band;
      const int i = k / (left + right), c = k - i * (left + right);
      gy = top + i;
      gx = c < left ? c : n1 - right + (c - left);
    }
  }
  __device__ __forceinline__ int index(int gy, int gx) const {              // (gy, gx) must be a frame cell
    if (gy < top) return gy * n1 + gx;
    if (gy >= n0 - bot) return (top + gy - (n0 - bot)) * n1 + gx;
    return n_band + (gy - top) * (left + right) + (gx < left ? gx : left + gx - (n1 - right));
  }
};

__device__ __forceinline__ float mw_edge_residual(const MatArgs& a, int gy, int gx) {
  const int n0 = a.n0, n1 = a.n1;
  float res = a.l1_fconst;
  if (a.l1_fbuf[0]) res += __ldg(a.l1_fbuf[0] + (size_t)gy * n1 + gx);
#pragma unroll 2
  for (int t = 0; t < a.n_lin; ++t) {
    const tdb200_mat_field& f = a.fld[a.lin_q[t]];
    float v;
    if (f.order == 0) {
      v = __ldg(a.u + (size_t)gy * n1 + gx);
    } else {
      float uv[2 * kMwMaxHw + 1], cv[2 * kMwMaxHw + 1];
#pragma unroll
      for (int m = -kMwMaxHw; m <= kMwMaxHw; ++m) {        // independent loads first
        const int yy = f.axis == 0 ? gy + m : gy, xx = f.axis == 0 ? gx : gx + m;
        const bool ok = m >= -f.half_width && m <= f.half_width && yy >= 0 && yy < n0 && xx >= 0 && xx < n1;
        uv[m + kMwMaxHw] = ok ? __ldg(a.u + (size_t)yy * n1 + xx) : 0.f;
        cv[m + kMwMaxHw] = ok ? band_coef(a.band, f, f.axis == 0 ? n0 : n1, f.axis == 0 ? gy : gx, m) : 0.f;
      }
      v = 0.f;
#pragma unroll
      for (int m = 0; m <= 2 * kMwMaxHw; ++m) v = fmaf(cv[m], uv[m], v);
    }
    res = fmaf(a.lin_c[t], v, res);
  }
  return res;
}

// phase A, run by the first `n_blocks` CTAs of the march launch
__device__ __forceinline__ double mw_edge_seeds(const MatArgs& a, int zy3, int zx3, float* __restrict__ es, int block, int n_blocks) {
  const MwFrame fr(a, zy3, zx3);
  const float scale2 = 2.f * a.eq_scale[0];
  double dacc = 0.0;
  for (int idx = block * (int)blockDim.x + (int)threadIdx.x; idx < fr.total; idx += n_blocks * (int)blockDim.x) {
    int gy, gx;
    fr.cell(idx, gy, gx);
    float seed = 0.f;
    if (!fr.regular(gy, gx)) {
      const float res = mw_edge_residual(a, gy, gx);
      if (gy >= a.row_lo && gy < a.row_hi) dacc += (double)res * (double)res;
      seed = scale2 * res;
    }
    es[idx] = seed;
  }
  return dacc;
}

// phase B
__global__ void __launch_bounds__(128) mat_march_edge_kernel(const MatArgs a, const int zy3, const int zx3,
                                                             const float* __restrict__ es) {
  const MwFrame fr(a, zy3, zx3);
  const int n0 = a.n0, n1 = a.n1;
  for (int idx = (int)(blockIdx.x * blockDim.x + threadIdx.x); idx < fr.total; idx += (int)(gridDim.x * blockDim.x)) {
    int gy, gx;
    fr.cell(idx, gy, gx);
    float g = 0.f;
#pragma unroll 2
    for (int t = 0; t < a.n_lin; ++t) {
      const tdb200_mat_field& f = a.fld[a.lin_q[t]];
      float sacc = 0.f;
      if (f.order == 0) {
        sacc = es[idx];                                      // zero for regular cells
      } else {
        float sv[2 * kMwMaxHw + 1], cv[2 * kMwMaxHw + 1];
#pragma unroll
        for (int m = -kMwMaxHw; m <= kMwMaxHw; ++m) {
          const int yy = f.axis == 0 ? gy + m : gy, xx = f.axis == 0 ? gx : gx + m;
          const bool ok = m >= -f.half_width && m <= f.half_width && yy >= 0 && yy < n0 && xx >= 0 && xx < n1 &&
                          !fr.regular(yy, xx);
          sv[m + kMwMaxHw] = ok ? __ldg(es + fr.index(yy, xx)) : 0.f;
          cv[m + kMwMaxHw] = ok ? band_coef(a.band, f, f.axis == 0 ? n0 : n1, f.axis == 0 ? yy : xx, -m) : 0.f;
        }
#pragma unroll
        for (int m = 0; m <= 2 * kMwMaxHw; ++m) sacc = fmaf(cv[m], sv[m], sacc);
      }
      g = fmaf(a.lin_c[t], sacc, g);
    }
    if (g != 0.f) a.grad[(size_t)gy * n1 + gx] += g;
  }
}

// One march item (a warp's column strip over a chunk of rows) of the register-marching kernel; returns the warp lane's
// share of the loss.  SKIP: gradient cells of the edge frame (rows < fzy from the top / bottom, columns < fzx from the
// left / right) are NOT stored - the edge CTAs of mat_march_fused_kernel own them.
// Rows [y0, y1) of row chunk `chunk`.  er == 0: uniform chunks of ch rows.  er > 0 (slab of a sharded grid): a short
// first and last chunk of er rows - the only ones whose stencils read halo rows, so the only warps that wait for the
// halo exchange, and they have a fraction of a regular chunk's work - and regular chunks of ch rows in between.
__host__ __device__ __forceinline__ void mw_chunk_rows(int n0, int ch, int er, int chunk, int& y0, int& y1) {
  if (er == 0) { y0 = chunk * ch; y1 = y0 + ch < n0 ? y0 + ch : n0; return; }
  const int nm = (n0 - 2 * er + ch - 1) / ch;              // regular chunks
  if (chunk == 0) { y0 = 0; y1 = er; }
  else if (chunk <= nm) { y0 = er + (chunk - 1) * ch; y1 = y0 + ch < n0 - er ? y0 + ch : n0 - er; }
  else { y0 = n0 - er; y1 = n0; }
}
__host__ __device__ __forceinline__ int mw_n_chunks(int n0, int ch, int er) {
  return er == 0 ? (n0 + ch - 1) / ch : 2 + (n0 - 2 * er + ch - 1) / ch;
}
template <int HY, int HX, unsigned MY, unsigned MX, int P, bool SKIP>
__device__ __forceinline__ double mw_march_item(const MatArgs& a, const int ch, const int n_strips, const int item,
                                                const int lane, const int fzy, const int fzx, const int er = 0) {
  constexpr int R = 2 * HY + 1 + P;                          // ring length = unroll factor of the row loop
  const int n0 = a.n0, n1 = a.n1;
  double dacc = 0.0;
  const int chunk = item / n_strips, strip = item - chunk * n_strips;
  int y0, y1;
  mw_chunk_rows(n0, ch, er, chunk, y0, y1);
  const int x = strip * kMwOutW - 4 + 4 * lane;            // column of this thread's first element
  const bool col_ok = x >= 0 && x < n1;                    // n1 % 4 == 0: the float4 is entirely inside or outside
  const bool own = col_ok && lane >= 1 && lane < 1 + kMwOutLanes;
  const bool frame_col = SKIP && (x < fzx || x >= n1 - fzx);     // fzx % 4 == 0: the float4 is entirely inside or outside the frame
  const bool f_ok = col_ok && a.l1_fbuf[0] != nullptr;
  const bool do_grad = a.grad != nullptr;
  const int zy = a.edge_y, zx = a.edge_x;
  // The stencil weights are read from the kernel parameters where they are used (constant-bank operands of the FFMAs,
  // no registers): that pays for one more row of register prefetch (P = 4).
  const float scale2 = 2.f * a.eq_scale[0];
  float sm2[4], lm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool reg = x + i >= zx && x + i < n1 - zx;       // regular column of the operators
    sm2[i] = reg ? scale2 : 0.f;
    lm[i] = (reg && own) ? 1.f : 0.f;
  }
  const int loss_lo = max(y0, a.row_lo), loss_hi = min(y1, a.row_hi);
  const int ybase = y0 - 2 * HY;                           // first row of u this warp loads
  const int load_hi = min(y1 + 2 * HY, n0), f_lo = max(y0 - HY, 0), f_hi = min(y1 + HY, n0);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 u[R], s[R], fr[R];
#pragma unroll
  for (int i = 0; i < R; ++i) u[i] = s[i] = fr[i] = zero4;
  const float* pu = a.u + (ptrdiff_t)ybase * n1 + x;                      // u row loaded at step j: ybase + j
  const float* pf = a.l1_fbuf[0] + (ptrdiff_t)(ybase - HY) * n1 + x;      // forcing row loaded at step j: ybase - HY + j
  float* pg = a.grad + (ptrdiff_t)(ybase - P - 2 * HY) * n1 + x;          // gradient row stored at step j
  const int steps = (y1 - y0) + 4 * HY + P;
  for (int jb = 0; jb < steps; jb += R) {
    float lacc = 0.f;
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      const int j = jb + jj;
      // (1) issue the loads of this step; they are consumed P steps from now
      {
        const int yl = ybase + j, yf = yl - HY;
        u[jj] = (col_ok && yl >= 0 && yl < load_hi) ? __ldg(reinterpret_cast<const float4*>(pu)) : zero4;
        fr[jj] = (f_ok && yf >= f_lo && yf < f_hi) ? __ldg(reinterpret_cast<const float4*>(pf)) : zero4;
        pu += n1; pf += n1;
      }
      // (2) residual row yr: centre = the u row loaded at step j - P - HY, forcing loaded at step j - P
      {
        const int yr = ybase + j - P - HY;
        const float4 c = u[(jj + 2 * R - P - HY) % R], fv = fr[(jj + R - P) % R];
        float r[4] = {a.l1_fconst + fv.x, a.l1_fconst + fv.y, a.l1_fconst + fv.z, a.l1_fconst + fv.w};
        r[0] = fmaf(a.cx_wc, c.x, r[0]); r[1] = fmaf(a.cx_wc, c.y, r[1]); r[2] = fmaf(a.cx_wc, c.z, r[2]); r[3] = fmaf(a.cx_wc, c.w, r[3]);
#pragma unroll
        for (int dy = -HY; dy <= HY; ++dy) {
          if (dy == 0 || !(MY & (1u << (dy + HY)))) continue;
          const float wgt = a.cx_wy[dy + HY];
          const float4 v = u[(jj + 2 * R - P - HY + dy) % R];
          r[0] = fmaf(wgt, v.x, r[0]); r[1] = fmaf(wgt, v.y, r[1]); r[2] = fmaf(wgt, v.z, r[2]); r[3] = fmaf(wgt, v.w, r[3]);
        }
        mw_xtaps<HX, MX, false>(c, a.cx_wx, r);
        const bool rowreg = yr >= zy && yr < n0 - zy;      // regular row of the operators (warp-uniform)
        if (rowreg && yr >= loss_lo && yr < loss_hi) {
#pragma unroll
          for (int i = 0; i < 4; ++i) lacc = fmaf(lm[i] * r[i], r[i], lacc);
        }
        s[jj] = rowreg ? make_float4(sm2[0] * r[0], sm2[1] * r[1], sm2[2] * r[2], sm2[3] * r[3]) : zero4;
      }
      // (3) gradient row yg = yr - HY: transposed stencil on the seed rows formed at steps j - 2 HY .. j
      if (do_grad) {
        const int yg = ybase + j - P - 2 * HY;
        const float4 c = s[(jj + R - HY) % R];
        float g[4] = {a.cx_wc * c.x, a.cx_wc * c.y, a.cx_wc * c.z, a.cx_wc * c.w};
#pragma unroll
        for (int dy = -HY; dy <= HY; ++dy) {
          if (dy == 0 || !(MY & (1u << (dy + HY)))) continue;
          const float wgt = a.cx_wy[dy + HY];
          const float4 v = s[(jj + 2 * R - HY - dy) % R];   // seed row yg - dy
          g[0] = fmaf(wgt, v.x, g[0]); g[1] = fmaf(wgt, v.y, g[1]); g[2] = fmaf(wgt, v.z, g[2]); g[3] = fmaf(wgt, v.w, g[3]);
        }
        mw_xtaps<HX, MX, true>(c, a.cx_wx, g);
        if (own && yg >= y0 && yg < y1 && !(SKIP && (frame_col || yg < fzy || yg >= n0 - fzy)))
          *reinterpret_cast<float4*>(pg) = make_float4(g[0], g[1], g[2], g[3]);
        pg += n1;
      }
    }
    dacc += (double)lacc;
  }
  return dacc;
}

template <int HY, int HX, unsigned MY, unsigned MX, int P>
__global__ void __launch_bounds__(kMwThreads, 2) mat_march_kernel(const MatArgs a, const int ch, const int n_strips,
                                                                  const int n_items, const int n_edge_blocks,
                                                                  float* __restrict__ edge_seeds, const int pdl_halo) {
  // pdl_halo: 0, or the rows of the short first / last chunk of a slab (see mw_chunk_rows)
  __shared__ double red[kMwWarps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = (int)blockIdx.x * kMwWarps + warp;
  const int first_edge_block = (int)gridDim.x - n_edge_blocks;
  double dacc = 0.0;
  // Slab of a sharded grid (pdl_halo): the launch is a programmatic dependent of the halo exchange in front of it
  // (peer_halo_kernel, csrc/peer.cu) - only the warps that READ halo rows wait for it, the rest of the slab is marched
  // while the rows cross NVLink.
  if (pdl_halo) {
    if ((int)blockIdx.x >= first_edge_block) pdl_wait();
    else if (item < n_items) {
      int y0, y1;
      mw_chunk_rows(a.n0, ch, pdl_halo, item / n_strips, y0, y1);
      if (y0 - 2 * HY < a.row_lo || y1 + 2 * HY > a.row_hi) pdl_wait();
    }
  }
  if ((int)blockIdx.x >= first_edge_block) {                 // phase A of the edge treatment: the trailing CTAs fill the
    dacc = mw_edge_seeds(a, a.edge_y + HY, a.edge_x + HX, edge_seeds, (int)blockIdx.x - first_edge_block, n_edge_blocks);   // tail
  } else if (item < n_items) {                               // warp-uniform
    dacc = mw_march_item<HY, HX, MY, MX, P, false>(a, ch, n_strips, item, lane, 0, 0, pdl_halo);
  }
  for (int o = 16; o; o >>= 1) dacc += __shfl_xor_sync(kFullMask, dacc, o);
  if (lane == 0) red[warp] = dacc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kMwWarps; ++w) t += red[w];
    a.part_loss[blockIdx.x] = t;
  }
}

constexpr int kMwEdgeBlocks = 296;                           // phase A CTAs (256 threads: one frame cell per thread at 4096^2;
                                                             // measured 148 -> 296: 62.7 -> 61.5 us per step)
// rows per chunk: every SM gets ~16 warps in a single wave (the y-halo of a chunk costs 4 HY extra row loads)
static int mat_march_chunk(const MatArgs& a, int n_sms) {
  const int n_strips = (a.n1 + kMwOutW - 1) / kMwOutW;
  const long long slots = (long long)n_sms * 2 * kMwWarps;
  long long ch = ((long long)a.n0 * n_strips + slots - 1) / slots;
  ch = (ch + 7) / 8 * 8;
  if (ch < 32) ch = 32;
  return (int)ch;
}
// slab of a sharded grid: rows of the short end chunks (halo rows + the rows whose stencils reach them), else 0
static int mat_march_edge_rows(const MatArgs& a, int hy, int ch) {
  if ((a.row_lo <= 0 && a.row_hi >= a.n0) || getenv("TDB200_NO_PDL")) return 0;
  const int halo = a.row_lo > a.n0 - a.row_hi ? a.row_lo : a.n0 - a.row_hi;
  const int er = (halo + 2 * hy + 7) / 8 * 8;
  return a.n0 - 2 * er >= ch ? er : 0;
}
static int mat_march_ctas(const MatArgs& a, int n_sms, int hy = 2) {         // loss partials written by the march launch
  const int ch = mat_march_chunk(a, n_sms), n_strips = (a.n1 + kMwOutW - 1) / kMwOutW;
  const int n_items = n_strips * mw_n_chunks(a.n0, ch, mat_march_edge_rows(a, hy, ch));
  return kMwEdgeBlocks + (n_items + kMwWarps - 1) / kMwWarps;
}
static size_t mat_march_frame_cells(const MatArgs& a, int hy, int hx) {
  const int zy3 = a.edge_y + hy, zx3 = a.edge_x + hx;
  const int top = zy3 < a.n0 ? zy3 : a.n0, bot = zy3 < a.n0 - top ? zy3 : a.n0 - top, mid = a.n0 - top - bot;
  const int left = zx3 < a.n1 ? zx3 : a.n1, right = zx3 < a.n1 - left ? zx3 : a.n1 - left;
  return (size_t)(top + bot) * a.n1 + (size_t)mid * (left + right);
}
template <int HY, int HX, unsigned MY, unsigned MX>
static cudaError_t launch_mat_march_t(const MatArgs& a, int n_sms, float* edge_seeds, cudaEvent_t after_stencil, bool main_only,
                                      cudaStream_t s) {
  const int ch = mat_march_chunk(a, n_sms), n_strips = (a.n1 + kMwOutW - 1) / kMwOutW;
  const int er = mat_march_edge_rows(a, HY, ch);
  const int n_items = n_strips * mw_n_chunks(a.n0, ch, er);
  const int grid = kMwEdgeBlocks + (n_items + kMwWarps - 1) / kMwWarps;
  // P = 5 rows of register prefetch: the deepest ring that does not spill at 128 registers (measured at 4096^2: P = 3
  // 65.6 us per step, 4: 63.4, 5: 62.6; P = 6 spills)
  cudaError_t e;
  if (er > 0) {
    e = launch_pdl(mat_march_kernel<HY, HX, MY, MX, 5>, dim3(grid), dim3(kMwThreads), 0, s, a, ch, n_strips, n_items,
                   (int)kMwEdgeBlocks, edge_seeds, er);
  } else {
    mat_march_kernel<HY, HX, MY, MX, 5><<<grid, kMwThreads, 0, s>>>(a, ch, n_strips, n_items, kMwEdgeBlocks, edge_seeds, 0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && after_stencil) e = cudaEventRecord(after_stencil, s);
  if (e != cudaSuccess || !a.grad || main_only) return e;
  mat_march_edge_kernel<<<kMwEdgeBlocks, 128, 0, s>>>(a, a.edge_y + HY, a.edge_x + HX, edge_seeds);
  return cudaGetLastError();
}
// instantiated for the cross shapes with reach <= 2 (wider stencils keep too many rows in registers)
#define TDB_MARCH_SHAPES(X) \
  X(2, 2, 0x11u, 0x11u) X(1, 2, 0x5u, 0x11u) X(2, 1, 0x11u, 0x5u) X(1, 1, 0x5u, 0x5u) X(2, 0, 0x11u, 0x0u) X(0, 2, 0x0u, 0x11u)
static bool mat_march_supported(int hy, int hx, unsigned my, unsigned mx) {
#define X(A, B, C, D) if (hy == A && hx == B && (my & ~C) == 0 && (mx & ~D) == 0) return true;
  TDB_MARCH_SHAPES(X)
#undef X
  return false;
}
static cudaError_t launch_mat_march(const MatArgs& a, int hy, int hx, unsigned my, unsigned mx, int n_sms, float* edge_seeds,
                                    cudaEvent_t after_stencil, bool main_only, cudaStream_t s) {
#define X(A, B, C, D) if (hy == A && hx == B && (my & ~C) == 0 && (mx & ~D) == 0) return launch_mat_march_t<A, B, C, D>(a, n_sms, edge_seeds, after_stencil, main_only, s);
  TDB_MARCH_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
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
      const bool core = ly >= hy && ly < hy + kMatTY && lx >= hx && lx < hx + kMatTX && gy >= a.row_lo && gy < a.row_hi;
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
      const bool core = ly >= hy && ly < hy + kMatTY && lx >= hx && lx < hx + kMatTX && gy >= a.row_lo && gy < a.row_hi;
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
  double* slot_sum;                        // [n_bc_slots]; zero on entry, reset by the finalizing block
  // finalize (run by the last block to finish): ordered reduction of the per-CTA loss partials + loss assembly
  const double* part_loss;
  int n_ctas, n_bc_slots;
  double n_cells;
  const double* slot_lambda;
  const double* slot_len;
  float* out;
  unsigned int* ticket;                    // zero on entry, reset by the finalizing block
  unsigned int* tile_ctr;                  // tile scheduler counter of the persistent kernel, reset likewise
  PeerLossArgs peer;                       // several ranks on one box (tdb200_mat_plan_set_peer): the finalizing block sums
                                           // the loss terms over the ranks through peer memory (world <= 1: off)
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

__device__ void mat_finalize_block(const double* __restrict__ part_loss, int n_ctas, int n_eq, double n_cells,
                                   double* __restrict__ bc_sum, int n_bc_slots, const double* __restrict__ slot_lambda,
                                   const double* __restrict__ slot_len, float* __restrict__ out);

__global__ void __launch_bounds__(128) mat_bc_kernel(const MatBcArgs a) {
  __shared__ double sh_slot[32];
  __shared__ unsigned int sh_ticket;
  if (threadIdx.x < 32) sh_slot[threadIdx.x] = 0.0;
  __syncthreads();
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
    {  // block-level accumulation: one shared-memory atomic per warp when the warp's rows share a slot
      const unsigned act = __activemask();
      double sq = (double)res * (double)res;
      const int slot0 = __shfl_sync(act, bc.slot, __ffs(act) - 1);
      if (act == 0xffffffffu && __all_sync(act, bc.slot == slot0)) {
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(act, sq, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sh_slot[slot0], sq);
      } else {
        atomicAdd(&sh_slot[bc.slot], sq);
      }
    }
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
  __syncthreads();
  if ((int)threadIdx.x < a.n_bc_slots && sh_slot[threadIdx.x] != 0.0) atomicAdd(a.slot_sum + threadIdx.x, sh_slot[threadIdx.x]);
  if (!a.out) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sh_ticket = atomicAdd(a.ticket, 1u);
  __syncthreads();
  if (sh_ticket != gridDim.x - 1) return;
  // last block: every other block's slot sums are visible
  __threadfence();
  if (threadIdx.x == 0) { *a.ticket = 0u; *a.tile_ctr = 0u; }
  mat_finalize_block(a.part_loss, a.n_ctas, a.n_eq, a.n_cells, a.slot_sum, a.n_bc_slots, a.slot_lambda, a.slot_len, a.out);
}

// Edge frame + boundary rows in ONE launch (register-marching plans whose boundary rows are all value rows on frame
// cells).  Blocks [0, n_frame_blocks): one thread per frame cell adds (a) phase B of the edge treatment - coef * seed
// over the special neighbours - and (b) the adjoint of every boundary row that touches the cell, found through a
// cell -> (condition, row, k) CSR built on the host, to the gradient the stencil kernel stored: one owner per gradient
// cell, fixed summation order, no atomics.  The remaining blocks reduce the boundary residuals into the loss slots.
// The last block to finish assembles the loss as mat_bc_kernel does.
struct MwEdgeBc {
  int zy3, zx3, n_frame_blocks;
  const float* es;                        // compact seeds written by phase A (previous launch)
  const int* csr_off;                     // [frame cells + 1]
  const int4* csr_ent;                    // {condition, row in the condition, k, unused}
};
__global__ void __launch_bounds__(128) mat_edge_bc_kernel(const MatBcArgs a, const MatArgs m, const MwEdgeBc e) {
  __shared__ double sh_slot[32];
  __shared__ unsigned int sh_ticket;
  if (threadIdx.x < 32) sh_slot[threadIdx.x] = 0.0;
  __syncthreads();
  const size_t N = (size_t)a.n0 * a.n1;
  auto row_residual = [&](const tdb200_mat_bc& bc, long long r) {
    float val = 0.f;
    for (int k = 0; k < bc.K; ++k) val += bc.sign[k] * __ldg(a.u + (size_t)bc.var * N + a.cells[bc.cell_off + r * bc.K + k]);
    return val - a.targets[bc.tgt_off + r];
  };
  if ((int)blockIdx.x < e.n_frame_blocks) {
    if (m.grad) {
      const MwFrame fr(m, e.zy3, e.zx3);
      const int n0 = m.n0, n1 = m.n1;
      for (int idx = (int)(blockIdx.x * blockDim.x + threadIdx.x); idx < fr.total; idx += e.n_frame_blocks * (int)blockDim.x) {
        int gy, gx;
        fr.cell(idx, gy, gx);
        float g = 0.f;
#pragma unroll 2
        for (int t = 0; t < m.n_lin; ++t) {
          const tdb200_mat_field& f = m.fld[m.lin_q[t]];
          float sacc = 0.f;
          if (f.order == 0) {
            sacc = __ldg(e.es + idx);                        // zero for regular cells
          } else {
            float sv[2 * kMwMaxHw + 1], cv[2 * kMwMaxHw + 1];
#pragma unroll
            for (int mm = -kMwMaxHw; mm <= kMwMaxHw; ++mm) {
              const int yy = f.axis == 0 ? gy + mm : gy, xx = f.axis == 0 ? gx : gx + mm;
              const bool ok = mm >= -f.half_width && mm <= f.half_width && yy >= 0 && yy < n0 && xx >= 0 && xx < n1 &&
                              !fr.regular(yy, xx);
              sv[mm + kMwMaxHw] = ok ? __ldg(e.es + fr.index(yy, xx)) : 0.f;
              cv[mm + kMwMaxHw] = ok ? band_coef(m.band, f, f.axis == 0 ? n0 : n1, f.axis == 0 ? yy : xx, -mm) : 0.f;
            }
#pragma unroll
            for (int mm = 0; mm <= 2 * kMwMaxHw; ++mm) sacc = fmaf(cv[mm], sv[mm], sacc);
          }
          g = fmaf(m.lin_c[t], sacc, g);
        }
        const int e0 = __ldg(e.csr_off + idx), e1 = __ldg(e.csr_off + idx + 1);
        for (int q = e0; q < e1; ++q) {                      // boundary rows that touch this cell, in row order
          const int4 en = __ldg(e.csr_ent + q);
          const tdb200_mat_bc bc = a.bcs[en.x];
          g = fmaf(2.f * a.slot_scale[bc.slot] * bc.sign[en.z], row_residual(bc, en.y), g);
        }
        if (g != 0.f) m.grad[(size_t)gy * n1 + gx] += g;
      }
    }
  } else {
    const long long total = a.bc_row_begin[a.n_bcs];
    const long long stride = (long long)((int)gridDim.x - e.n_frame_blocks) * blockDim.x;
    for (long long row = (long long)((int)blockIdx.x - e.n_frame_blocks) * blockDim.x + threadIdx.x; row < total; row += stride) {
      int bi = 0;
      while (row >= a.bc_row_begin[bi + 1]) ++bi;
      const tdb200_mat_bc bc = a.bcs[bi];
      const float res = row_residual(bc, row - a.bc_row_begin[bi]);
      const unsigned act = __activemask();
      double sq = (double)res * (double)res;
      const int slot0 = __shfl_sync(act, bc.slot, __ffs(act) - 1);
      if (act == 0xffffffffu && __all_sync(act, bc.slot == slot0)) {      // one shared-memory atomic per warp
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(act, sq, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sh_slot[slot0], sq);
      } else {
        atomicAdd(&sh_slot[bc.slot], sq);
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < a.n_bc_slots && sh_slot[threadIdx.x] != 0.0) atomicAdd(a.slot_sum + threadIdx.x, sh_slot[threadIdx.x]);
  if (!a.out) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sh_ticket = atomicAdd(a.ticket, 1u);
  __syncthreads();
  if (sh_ticket != gridDim.x - 1) return;
  __threadfence();
  if (threadIdx.x == 0) { *a.ticket = 0u; *a.tile_ctr = 0u; }
  mat_finalize_block(a.part_loss, a.n_ctas, a.n_eq, a.n_cells, a.slot_sum, a.n_bc_slots, a.slot_lambda, a.slot_len, a.out);
  if (a.peer.world > 1) {                  // (block-uniform) loss terms of all ranks, summed in rank order
    __syncthreads();
    peer_loss_block(a.peer.blocks, a.peer.rank, a.peer.world, a.peer.step_dev, a.out, 2 + a.n_eq + a.n_bc_slots);
  }
}

// ordered reduction of the per-CTA loss partials + loss assembly, by one block of any size <= 256
__device__ void mat_finalize_block(const double* __restrict__ part_loss, int n_ctas, int n_eq, double n_cells,
                                   double* __restrict__ bc_sum, int n_bc_slots, const double* __restrict__ slot_lambda,
                                   const double* __restrict__ slot_len, float* __restrict__ out) {
  __shared__ double sh[256];
  __shared__ double mse[32];
  for (int e = 0; e < n_eq; ++e) {
    double s = 0.0;
    for (int c = threadIdx.x; c < n_ctas; c += blockDim.x) s += part_loss[(size_t)c * n_eq + e];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += sh[w];
      mse[e] = t / slot_len[e];          // global cell count (slab decomposition: partial sums add across ranks)
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double loss = 0.0, lossn = 0.0;
    for (int e = 0; e < n_eq; ++e) { out[2 + e] = (float)mse[e]; loss += slot_lambda[e] * mse[e]; lossn += mse[e]; }
    for (int s = 0; s < n_bc_slots; ++s) {
      const double m = *reinterpret_cast<volatile double*>(bc_sum + s) / slot_len[n_eq + s];
      out[2 + n_eq + s] = (float)m;
      loss += slot_lambda[n_eq + s] * m;
      lossn += m;
      bc_sum[s] = 0.0;                                       // ready for the next call
    }
    out[0] = (float)loss;
    out[1] = (float)lossn;
  }
}


// ------------------------------------------------------------------------------------------------------
// EXPERIMENT (opt-in with TDB200_MAT_FUSED=1; measured 68.7 us per step at 4096^2 against 59.7 us of the two-launch
// schedule: the edge phases are multi-microsecond chains of dependent loads wherever they run, and the union of the
// roles spills in the marching loop at 128 registers).
// The whole mat-mode step in ONE launch (plans of mat_edge_bc_kernel: one linear constant-coefficient equation, every
// boundary row a value row on a frame cell - BASELINE config 4).  The step used to be march launch (with phase A in its
// tail) -> edge / boundary launch: 36 us of streaming followed by ~20 us of latency chains of two tiny grids.  Here the
// first `n_edge_blocks` CTAs of the launch are EDGE CTAs, resident next to the marching CTAs from the start:
//   A'  seeds 2 lambda / N * residual of every cell of the EXTENDED frame (frame + stencil reach), from u and the banded
//       operators; loss of the cells with special rows (the marching warps count the regular ones);
//   --  barrier among the edge CTAs (they are the first CTAs of the grid: co-resident, a spin on a counter is safe);
//   B'  every frame cell gathers coef * seed over ALL its neighbours (regular rows included) plus the adjoint of the
//       boundary rows that touch it (cell -> row CSR) and STORES its gradient: the frame cells belong to the edge CTAs
//       alone, the marching warps skip them (mw_march_item<SKIP>), so nothing orders the two roles;
//   C   boundary residuals -> loss slots.
// The last CTA of the grid to finish assembles the loss.  One owner per gradient cell, fixed summation orders:
// bit-reproducible like the two-launch schedule.
template <int HY, int HX, unsigned MY, unsigned MX, int P>
__global__ void __launch_bounds__(kMwThreads, 2) mat_march_fused_kernel(const MatArgs a, const int ch, const int n_strips,
                                                                        const int n_items, const int n_edge_blocks,
                                                                        float* __restrict__ es, const MatBcArgs b,
                                                                        const MwEdgeBc e, unsigned int* edge_sync) {
  __shared__ double red[kMwWarps];
  __shared__ double sh_slot[32];
  __shared__ unsigned int sh_ticket;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = a.n0, n1 = a.n1;
  double dacc = 0.0;
  if (tid < 32) sh_slot[tid] = 0.0;
  __syncthreads();
  {                                                                                           // A': every CTA takes a share
    const MwFrame frx(a, e.zy3 + HY, e.zx3 + HX);
    const float scale2 = 2.f * a.eq_scale[0];
    for (int idx = (int)blockIdx.x * (int)blockDim.x + tid; idx < frx.total; idx += (int)gridDim.x * (int)blockDim.x) {
      int gy, gx;
      frx.cell(idx, gy, gx);
      const float res = mw_edge_residual(a, gy, gx);
      if (!frx.regular(gy, gx) && gy >= a.row_lo && gy < a.row_hi) dacc += (double)res * (double)res;
      es[idx] = scale2 * res;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(edge_sync, 1u);
  }
  {                                                                                           // the marching work
    const int item = (int)blockIdx.x * kMwWarps + warp;
    if (item < n_items) dacc += mw_march_item<HY, HX, MY, MX, P, true>(a, ch, n_strips, item, lane, e.zy3, e.zx3);
  }
  {                                                                                           // B' + C: every CTA takes a share
    const MwFrame frx(a, e.zy3 + HY, e.zx3 + HX), fr(a, e.zy3, e.zx3);
    const int stride = (int)gridDim.x * (int)blockDim.x, first = (int)blockIdx.x * (int)blockDim.x + tid;
    if (tid == 0)                                    // every CTA of the grid is resident (checked by the host) and did A'
      while (atomicAdd(edge_sync, 0u) < gridDim.x) __nanosleep(100);   // before its march: no waiting in practice
    __syncthreads();
    __threadfence();
    const size_t N = (size_t)n0 * n1;
    auto row_residual = [&](const tdb200_mat_bc& bc, long long r) {
      float val = 0.f;
      for (int k = 0; k < bc.K; ++k) val += bc.sign[k] * __ldg(b.u + (size_t)bc.var * N + b.cells[bc.cell_off + r * bc.K + k]);
      return val - b.targets[bc.tgt_off + r];
    };
    for (int idx = first; idx < fr.total; idx += stride) {                                    // B'
      int gy, gx;
      fr.cell(idx, gy, gx);
      float g = 0.f;
#pragma unroll 2
      for (int t = 0; t < a.n_lin; ++t) {
        const tdb200_mat_field& f = a.fld[a.lin_q[t]];
        float sacc = 0.f;
        if (f.order == 0) {
          sacc = __ldcg(es + frx.index(gy, gx));
        } else {
          float sv[2 * kMwMaxHw + 1], cv[2 * kMwMaxHw + 1];
#pragma unroll
          for (int mm = -kMwMaxHw; mm <= kMwMaxHw; ++mm) {
            const int yy = f.axis == 0 ? gy + mm : gy, xx = f.axis == 0 ? gx : gx + mm;
            const bool ok = mm >= -f.half_width && mm <= f.half_width && yy >= 0 && yy < n0 && xx >= 0 && xx < n1;
            sv[mm + kMwMaxHw] = ok ? __ldcg(es + frx.index(yy, xx)) : 0.f;
            cv[mm + kMwMaxHw] = ok ? band_coef(a.band, f, f.axis == 0 ? n0 : n1, f.axis == 0 ? yy : xx, -mm) : 0.f;
          }
#pragma unroll
          for (int mm = 0; mm <= 2 * kMwMaxHw; ++mm) sacc = fmaf(cv[mm], sv[mm], sacc);
        }
        g = fmaf(a.lin_c[t], sacc, g);
      }
      const int e0 = __ldg(e.csr_off + idx), e1 = __ldg(e.csr_off + idx + 1);
      for (int q = e0; q < e1; ++q) {                      // boundary rows that touch this cell, in row order
        const int4 en = __ldg(e.csr_ent + q);
        const tdb200_mat_bc bc = b.bcs[en.x];
        g = fmaf(2.f * b.slot_scale[bc.slot] * bc.sign[en.z], row_residual(bc, en.y), g);
      }
      a.grad[(size_t)gy * n1 + gx] = g;
    }
    const long long total = b.bc_row_begin[b.n_bcs];                                          // C
    for (long long row = first; row < total; row += stride) {
      int bi = 0;
      while (row >= b.bc_row_begin[bi + 1]) ++bi;
      const tdb200_mat_bc bc = b.bcs[bi];
      const float res = row_residual(bc, row - b.bc_row_begin[bi]);
      const unsigned act = __activemask();
      double sq = (double)res * (double)res;
      const int slot0 = __shfl_sync(act, bc.slot, __ffs(act) - 1);
      if (act == 0xffffffffu && __all_sync(act, bc.slot == slot0)) {      // one shared-memory atomic per warp
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(act, sq, o);
        if (lane == 0) atomicAdd(&sh_slot[slot0], sq);
      } else {
        atomicAdd(&sh_slot[bc.slot], sq);
      }
    }
  }
  for (int o = 16; o; o >>= 1) dacc += __shfl_xor_sync(kFullMask, dacc, o);
  if (lane == 0) red[warp] = dacc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kMwWarps; ++w) t += red[w];
    a.part_loss[blockIdx.x] = t;
  }
  if (tid < b.n_bc_slots && sh_slot[tid] != 0.0) atomicAdd(b.slot_sum + tid, sh_slot[tid]);
  __threadfence();
  __syncthreads();
  if (tid == 0) sh_ticket = atomicAdd(b.ticket, 1u);
  __syncthreads();
  if (sh_ticket != gridDim.x - 1) return;
  __threadfence();
  if (tid == 0) { *b.ticket = 0u; *edge_sync = 0u; }
  mat_finalize_block(a.part_loss, (int)gridDim.x, b.n_eq, b.n_cells, b.slot_sum, b.n_bc_slots, b.slot_lambda, b.slot_len, b.out);
}

constexpr int kMwFusedEdgeBlocks = 16;
static int mat_march_fused_ctas(const MatArgs& a, int n_sms);        // 280 marching CTAs + 16 edge CTAs = 296 = 2 CTAs on each of the 148 SMs at 4096^2
template <int HY, int HX, unsigned MY, unsigned MX>
static cudaError_t launch_mat_march_fused_t(const MatArgs& a, int n_sms, float* es, const MatBcArgs& b, const MwEdgeBc& e,
                                            unsigned int* edge_sync, int* n_ctas, cudaStream_t s) {
  const int ch = mat_march_chunk(a, n_sms), n_strips = (a.n1 + kMwOutW - 1) / kMwOutW;
  const int n_items = n_strips * ((a.n0 + ch - 1) / ch);
  const int grid = kMwFusedEdgeBlocks + (n_items + kMwWarps - 1) / kMwWarps;
  *n_ctas = grid;
  mat_march_fused_kernel<HY, HX, MY, MX, 5><<<grid, kMwThreads, 0, s>>>(a, ch, n_strips, n_items, kMwFusedEdgeBlocks, es, b, e, edge_sync);
  return cudaGetLastError();
}
static int mat_march_fused_ctas(const MatArgs& a, int n_sms) {
  const int ch = mat_march_chunk(a, n_sms), n_strips = (a.n1 + kMwOutW - 1) / kMwOutW;
  return kMwFusedEdgeBlocks + (n_strips * ((a.n0 + ch - 1) / ch) + kMwWarps - 1) / kMwWarps;
}
static cudaError_t launch_mat_march_fused(const MatArgs& a, int hy, int hx, unsigned my, unsigned mx, int n_sms, float* es,
                                          const MatBcArgs& b, const MwEdgeBc& e, unsigned int* edge_sync, int* n_ctas,
                                          cudaStream_t s) {
#define X(A, B, C, D) if (hy == A && hx == B && (my & ~C) == 0 && (mx & ~D) == 0) return launch_mat_march_fused_t<A, B, C, D>(a, n_sms, es, b, e, edge_sync, n_ctas, s);
  TDB_MARCH_SHAPES(X)
#undef X
  return cudaErrorInvalidValue;
}

__global__ void mat_finalize_kernel(const double* __restrict__ part_loss, int n_ctas, int n_eq, double n_cells,
                                    double* __restrict__ bc_sum, int n_bc_slots,
                                    const double* __restrict__ slot_lambda, const double* __restrict__ slot_len,
                                    float* __restrict__ out, unsigned int* tile_ctr) {
  if (threadIdx.x == 0) *tile_ctr = 0u;
  mat_finalize_block(part_loss, n_ctas, n_eq, n_cells, bc_sum, n_bc_slots, slot_lambda, slot_len, out);
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

struct tdb200_peer;
extern "C" int tdb200_peer_export_(tdb200_peer* p, tdb::PeerLossArgs* out);

struct tdb200_mat_plan {
  tdb200_peer* peer = nullptr;             // borrowed (tdb200_mat_plan_set_peer)
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
  long long l1_fbuf_off[2] = {-1, -1};     // lin1 kernel: forcing buffer offsets into the coefficient arena
  int l1_n_fbuf = 0;
  bool l1_regs = false;
  double* d_bc_sum = nullptr;
  unsigned int* d_ticket = nullptr;
  int cx_hy = 0, cx_hx = 0;                // cross kernel: reach and non-zero offset masks of the composite stencil
  unsigned cx_my = 0, cx_mx = 0;
  bool cross = false;
  bool tma = false;                        // persistent TMA variant of the cross kernel (single forcing buffer)
  bool march = false;                      // register-marching variant (reach <= 2, single forcing buffer)
  float* d_edge_seed = nullptr;            // march kernel: compact seeds of the frame cells with special stencil rows
  bool edge_bc = false;                    // march plans: edge frame + boundary rows in one launch (value rows on frame cells)
  int* d_csr_off = nullptr;                // frame cell -> boundary rows touching it
  int4* d_csr_ent = nullptr;
  int n_frame_blocks = 0;
  bool timing = false;                     // measurement aid: CUDA events around the stencil kernel launch
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  CUtensorMap map_u{}, map_f{};
  const float* map_u_ptr = nullptr;
  const float* map_f_ptr = nullptr;
  int n_sms = 148;
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
    int n_fbuf = 0;
    float fconst = 0.f;
    for (int t = 0; t < n_frc && a.linear; ++t) {
      if (a.frc_buf[t] >= 0) { if (n_fbuf < 2) p->l1_fbuf_off[n_fbuf] = a.frc_buf[t]; ++n_fbuf; }
      else fconst += a.frc_const[t];
    }
    a.l1_fconst = fconst;
    p->l1_n_fbuf = n_fbuf;
    a.lin1 = (a.linear && desc->n_eq == 1 && desc->n_var == 1 && n_fbuf <= 2 && hy <= tdb::kL1MaxH &&
              hx <= tdb::kL1MaxH) ? 1 : 0;
    p->l1_regs = n_taps <= 16;                            // the register-tap kernel holds at most 16 taps
    if (getenv("TDB200_MAT_NO_LIN1")) a.lin1 = 0;        // debugging aid: force the generic kernel
    if (a.lin1) {                                         // composite cross stencil: weights by offset
      for (int i = 0; i < 9; ++i) a.cx_wy[i] = a.cx_wx[i] = 0.f;
      a.cx_wc = 0.f;
      unsigned my = 0, mx = 0;
      for (int t = 0; t < n_taps; ++t) {
        const int m = a.tap_m[t];
        if (m == 0) a.cx_wc += a.tap_w[t];
        else if (a.tap_axis[t] == 0) { a.cx_wy[m + hy] += a.tap_w[t]; my |= 1u << (m + hy); }
        else { a.cx_wx[m + hx] += a.tap_w[t]; mx |= 1u << (m + hx); }
      }
      p->cx_hy = hy; p->cx_hx = hx; p->cx_my = my; p->cx_mx = mx;
      p->cross = tdb::mat_cross_supported(hy, hx, my, mx) && desc->n1 % 4 == 0 && !getenv("TDB200_MAT_NO_CROSS");
      if (!p->cross && !p->l1_regs) a.lin1 = 0;           // neither specialised kernel applies
      p->tma = p->cross && n_fbuf <= 1 && !getenv("TDB200_MAT_NO_TMA");
      p->march = p->cross && n_fbuf <= 1 && tdb::mat_march_supported(hy, hx, my, mx) && !getenv("TDB200_MAT_NO_MARCH") &&
                 (long long)(desc->n0 + 64) * desc->n1 < (1ll << 31);      // the kernel uses 32-bit element offsets
      cudaDeviceGetAttribute(&p->n_sms, cudaDevAttrMultiProcessorCount, device);
    }
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
  {
    int l1 = a.lin1 ? tdb::mat_lin1_ctas(a) : 0;         // (the cross kernel uses the same tiling)
    if (p->march) {
      // (+ the CTAs of the two short end chunks a row window adds later, see mw_chunk_rows)
      const int m = tdb::mat_march_ctas(a, p->n_sms) + 2 * (((a.n1 + tdb::kMwOutW - 1) / tdb::kMwOutW + tdb::kMwWarps - 1) / tdb::kMwWarps) + 2;
      l1 = m > l1 ? m : l1;
      MCU(cudaMalloc(&p->d_edge_seed, sizeof(float) * (tdb::mat_march_frame_cells(a, 2 * hy, 2 * hx) + 1)));   // extended frame
    }
    MCU(cudaMalloc(&p->d_part_loss, sizeof(double) * (size_t)(p->n_ctas > l1 ? p->n_ctas : l1) * desc->n_eq));
  }
  a.band = p->d_band; a.terms = p->d_terms; a.factors = p->d_factors; a.part_loss = p->d_part_loss;
  a.row_lo = 0; a.row_hi = desc->n0;
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
  MCU(cudaMemset(p->d_bc_sum, 0, sizeof(double) * (p->n_bc_slots > 0 ? p->n_bc_slots : 1)));
  if (!p->d_ticket) MCU(cudaMalloc(&p->d_ticket, 4 * sizeof(unsigned int)));   // [finalize ticket, tile counter, edge barrier]
  MCU(cudaMemset(p->d_ticket, 0, 4 * sizeof(unsigned int)));
  for (int e = 0; e < n_eq; ++e) p->args.eq_scale[e] = (float)(slot_lambda[e] / slot_len[e]);
  {  // register-marching plans: can the edge frame and the boundary rows share one launch?  (every row a value row -
     // Dirichlet / periodic values - and every cell it touches a frame cell: then the adjoint of the rows is gathered per
     // frame cell through a cell -> row CSR instead of scattered with atomics after the edge launch)
    p->edge_bc = false;
    cudaFree(p->d_csr_off); cudaFree(p->d_csr_ent);
    p->d_csr_off = nullptr; p->d_csr_ent = nullptr;
    bool ok = p->march && p->args.lin1 && p->desc.n_var == 1 && !getenv("TDB200_MAT_NO_EDGE_BC");
    long long n_cells_total = 0;
    for (int i = 0; i < n_bcs && ok; ++i) {
      ok = bcs[i].term_begin == bcs[i].term_end && bcs[i].var == 0;
      n_cells_total += bcs[i].n_rows * bcs[i].K;
    }
    if (ok && n_cells_total > 0 && n_cells_total < (1ll << 24)) {
      const tdb::MatArgs& a = p->args;
      const int n0 = a.n0, n1 = a.n1, zy3 = a.edge_y + p->cx_hy, zx3 = a.edge_x + p->cx_hx;
      const int top = zy3 < n0 ? zy3 : n0, bot = zy3 < n0 - top ? zy3 : n0 - top, mid = n0 - top - bot;
      const int left = zx3 < n1 ? zx3 : n1, right = zx3 < n1 - left ? zx3 : n1 - left;
      const long long n_band = (long long)(top + bot) * n1, total = n_band + (long long)mid * (left + right);
      auto frame_index = [&](int gy, int gx) -> long long {
        if (gy < top) return (long long)gy * n1 + gx;
        if (gy >= n0 - bot) return (long long)(top + gy - (n0 - bot)) * n1 + gx;
        if (gx < left) return n_band + (long long)(gy - top) * (left + right) + gx;
        if (gx >= n1 - right) return n_band + (long long)(gy - top) * (left + right) + left + gx - (n1 - right);
        return -1;
      };
      std::vector<int> cells((size_t)n_cells_total);
      // the cells of condition i start at bcs[i].cell_off: copy the used range of cells_dev
      long long lo = -1, hi = 0;
      for (int i = 0; i < n_bcs; ++i) {
        if (lo < 0 || bcs[i].cell_off < lo) lo = bcs[i].cell_off;
        const long long end = bcs[i].cell_off + bcs[i].n_rows * bcs[i].K;
        hi = end > hi ? end : hi;
      }
      std::vector<int> host((size_t)(hi - lo));
      MCU(cudaMemcpy(host.data(), cells_dev + lo, sizeof(int) * host.size(), cudaMemcpyDeviceToHost));
      std::vector<int> count((size_t)total + 1, 0);
      for (int i = 0; i < n_bcs && ok; ++i)
        for (long long r = 0; r < bcs[i].n_rows && ok; ++r)
          for (int k = 0; k < bcs[i].K; ++k) {
            const int cell = host[(size_t)(bcs[i].cell_off - lo + r * bcs[i].K + k)];
            const long long fi = (cell >= 0 && cell < n0 * n1) ? frame_index(cell / n1, cell % n1) : -1;
            if (fi < 0) { ok = false; break; }
            ++count[(size_t)fi + 1];
          }
      if (ok) {
        for (size_t c = 0; c < (size_t)total; ++c) count[c + 1] += count[c];
        std::vector<int4> ent((size_t)count[(size_t)total]);
        std::vector<int> fill(count.begin(), count.end() - 1);
        for (int i = 0; i < n_bcs; ++i)                       // condition-major, row-major: a fixed summation order per cell
          for (long long r = 0; r < bcs[i].n_rows; ++r)
            for (int k = 0; k < bcs[i].K; ++k) {
              const int cell = host[(size_t)(bcs[i].cell_off - lo + r * bcs[i].K + k)];
              const long long fi = frame_index(cell / n1, cell % n1);
              ent[(size_t)fill[(size_t)fi]++] = make_int4(i, (int)r, k, 0);
            }
        MCU(cudaMalloc(&p->d_csr_off, sizeof(int) * count.size()));
        MCU(cudaMemcpy(p->d_csr_off, count.data(), sizeof(int) * count.size(), cudaMemcpyHostToDevice));
        MCU(cudaMalloc(&p->d_csr_ent, sizeof(int4) * (ent.size() > 0 ? ent.size() : 1)));
        if (!ent.empty()) MCU(cudaMemcpy(p->d_csr_ent, ent.data(), sizeof(int4) * ent.size(), cudaMemcpyHostToDevice));
        const long long fb = (total + 127) / 128;
        p->n_frame_blocks = (int)(fb < 1 ? 1 : fb > 592 ? 592 : fb);
        p->edge_bc = true;
      }
    }
  }
  tdb::MatBcArgs& b = p->bc;
  b.bcs = p->d_bcs; b.n_bcs = n_bcs; b.bc_row_begin = p->d_bc_row_begin; b.cells = cells_dev; b.targets = targets_dev;
  b.slot_scale = p->d_slot_scale; b.slot_sum = p->d_bc_sum;
  p->bcs_set = true;
  return TDB200_OK;
}

// the residual (stencil) launch(es) of one call; *n_ctas = number of loss partials they write
static int mat_stencil(tdb200_mat_plan* p, tdb::MatArgs& a, const float* u, float* grad, float* op_out, int* n_ctas_out,
                       cudaEvent_t after_stencil, bool main_only, cudaStream_t s, bool* ran_march = nullptr) {
  if (ran_march) *ran_march = false;
  int n_ctas = p->n_ctas;
  bool ev1_done = false;
  if (a.lin1 && !op_out) {
    // specialised kernels (loss + gradient, or loss only); per-cell operator values go through the generic kernel
    for (int i = 0; i < 2; ++i) a.l1_fbuf[i] = i < p->l1_n_fbuf ? a.coeffs + p->l1_fbuf_off[i] : nullptr;
    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec_ok = aligned16(u) && aligned16(grad) && aligned16(a.l1_fbuf[0]) && aligned16(a.l1_fbuf[1]);
    if (p->cross && !vec_ok && !p->l1_regs) return mat_invalid("mat-mode tensors must be 16-byte aligned");
    bool tma = p->tma && vec_ok && !p->march;
    if (tma) {                                             // tensor maps: re-encoded only when a pointer changes
      const int uy = tdb::kCxTY + 4 * p->cx_hy, ry = tdb::kCxTY + 2 * p->cx_hy;
      if (p->map_u_ptr != u) {
        tma = tdb::make_map_2d(&p->map_u, u, a.n0, a.n1, uy, tdb::kCxPU);
        p->map_u_ptr = tma ? u : nullptr;
      }
      if (tma && a.l1_fbuf[0] && p->map_f_ptr != a.l1_fbuf[0]) {
        tma = tdb::make_map_2d(&p->map_f, a.l1_fbuf[0], a.n0, a.n1, ry, tdb::kCxPR);
        p->map_f_ptr = tma ? a.l1_fbuf[0] : nullptr;
      }
    }
    if (p->march && vec_ok) {
      MCU(tdb::launch_mat_march(a, p->cx_hy, p->cx_hx, p->cx_my, p->cx_mx, p->n_sms, p->d_edge_seed, after_stencil,
                                main_only || p->edge_bc, s));
      if (ran_march) *ran_march = true;
      ev1_done = true;
      n_ctas = tdb::mat_march_ctas(a, p->n_sms, p->cx_hy);
    } else if (tma) {
      MCU(tdb::launch_mat_cross_tma(a, p->cx_hy, p->cx_hx, p->cx_my, p->cx_mx, p->map_u, a.l1_fbuf[0] ? p->map_f : p->map_u,
                                    p->n_sms, &n_ctas, s));
    } else if (p->cross && vec_ok) {
      MCU(tdb::launch_mat_cross(a, p->cx_hy, p->cx_hx, p->cx_my, p->cx_mx, s));
      n_ctas = tdb::mat_cross_ctas(a);
    } else {
      MCU(tdb::launch_mat_lin1(a, s));
      n_ctas = tdb::mat_lin1_ctas(a);
    }
  } else {
    dim3 grid(a.tiles_x, a.tiles_y);
    tdb::mat_residual_adjoint_kernel<<<grid, tdb::kMatThreads, p->smem, s>>>(a);
    MCU(cudaGetLastError());
  }
  if (after_stencil && !ev1_done) MCU(cudaEventRecord(after_stencil, s));
  *n_ctas_out = n_ctas;
  return TDB200_OK;
}

// single-launch step: a march plan whose boundary rows are value rows on frame cells, loss + gradient call, 16-byte
// aligned tensors, frame columns a multiple of 4 wide (the marching warps skip whole float4s)
static bool mat_fused_ok(const tdb200_mat_plan* p, const tdb::MatArgs& a, const float* u, const float* grad, const float* op_out,
                         const float* bval_out) {
  if (!(p->march && p->edge_bc && p->args.lin1 && grad && !op_out && !bval_out && p->n_bc_rows > 0)) return false;
  if (!getenv("TDB200_MAT_FUSED")) return false;          // measured slower than the two-launch schedule (68.7 vs 59.7 us): opt-in
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const float* fb = p->l1_n_fbuf > 0 ? p->args.coeffs + p->l1_fbuf_off[0] : nullptr;
  if (!aligned16(u) || !aligned16(grad) || !aligned16(fb)) return false;
  const int zy3 = a.edge_y + p->cx_hy, zx3 = a.edge_x + p->cx_hx;
  // the edge CTAs wait for every CTA of the grid: all of them must be resident at once (2 CTAs of 256 threads per SM)
  if (tdb::mat_march_fused_ctas(a, p->n_sms) > 2 * p->n_sms) return false;
  return zx3 % 4 == 0 && a.n1 % 4 == 0 && 2 * (zx3 + p->cx_hx) < a.n1 && 2 * (zy3 + p->cx_hy) < a.n0;
}

static int mat_run(tdb200_mat_plan* p, const float* u, float* grad, float* op_out, float* bval_out, float* out,
                   void* stream) {
  if (!p || !u || !out) return mat_invalid("null argument");
  if (!p->bcs_set) return mat_invalid("tdb200_mat_plan_set_bcs was not called");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MCU(cudaSetDevice(p->device));
  tdb::MatArgs a = p->args;
  a.u = u; a.grad = grad; a.op_out = op_out; a.tile_ctr = p->d_ticket + 1;
  int n_ctas = p->n_ctas;
  if (p->timing) MCU(cudaEventRecord(p->ev0, s));
  bool ran_march = false;
  if (mat_fused_ok(p, a, u, grad, op_out, bval_out)) {
    // the whole step in one launch: marching CTAs + edge CTAs (seeds of the extended frame, frame gradient, boundary
    // rows) + loss assembly by the last CTA
    for (int i = 0; i < 2; ++i) a.l1_fbuf[i] = i < p->l1_n_fbuf ? a.coeffs + p->l1_fbuf_off[i] : nullptr;
    tdb::MatBcArgs b = p->bc;
    b.u = u; b.grad = grad; b.bval_out = nullptr;
    b.part_loss = p->d_part_loss; b.n_bc_slots = p->n_bc_slots;
    b.n_cells = (double)p->desc.n0 * (double)p->desc.n1;
    b.slot_lambda = p->d_slot_lambda; b.slot_len = p->d_slot_len; b.out = out; b.ticket = p->d_ticket; b.tile_ctr = p->d_ticket + 1;
    tdb::MwEdgeBc eb{};
    eb.zy3 = a.edge_y + p->cx_hy; eb.zx3 = a.edge_x + p->cx_hx; eb.n_frame_blocks = 0;
    eb.es = p->d_edge_seed; eb.csr_off = p->d_csr_off; eb.csr_ent = p->d_csr_ent;
    MCU(tdb::launch_mat_march_fused(a, p->cx_hy, p->cx_hx, p->cx_my, p->cx_mx, p->n_sms, p->d_edge_seed, b, eb, p->d_ticket + 2,
                                    &n_ctas, s));
    b.n_ctas = n_ctas;
    if (p->timing) MCU(cudaEventRecord(p->ev1, s));
    return TDB200_OK;
  }
  {
    const int rc = mat_stencil(p, a, u, grad, op_out, &n_ctas, p->timing ? p->ev1 : nullptr, false, s, &ran_march);
    if (rc != TDB200_OK) return rc;
  }
  // boundary rows; the last block to finish reduces the loss partials and assembles the loss (and re-zeroes the
  // slot sums and its ticket for the next call)
  const bool merged = p->edge_bc && ran_march && p->n_bc_rows > 0;
  if (p->n_bc_rows > 0) {
    tdb::MatBcArgs b = p->bc;
    b.u = u; b.grad = grad; b.bval_out = bval_out;
    b.part_loss = p->d_part_loss; b.n_ctas = n_ctas; b.n_bc_slots = p->n_bc_slots;
    b.n_cells = (double)p->desc.n0 * (double)p->desc.n1;
    b.slot_lambda = p->d_slot_lambda; b.slot_len = p->d_slot_len; b.out = out; b.ticket = p->d_ticket; b.tile_ctr = p->d_ticket + 1;
    b.peer.world = 0;
    const int blocks = (int)((p->n_bc_rows + 127) / 128);
    if (merged) {
      if (p->peer && out) { const int rc = tdb200_peer_export_(p->peer, &b.peer); if (rc != TDB200_OK) return rc; }
      tdb::MwEdgeBc eb{};
      eb.zy3 = a.edge_y + p->cx_hy; eb.zx3 = a.edge_x + p->cx_hx; eb.n_frame_blocks = p->n_frame_blocks;
      eb.es = p->d_edge_seed; eb.csr_off = p->d_csr_off; eb.csr_ent = p->d_csr_ent;
      tdb::mat_edge_bc_kernel<<<p->n_frame_blocks + (blocks < 1184 ? blocks : 1184), 128, 0, s>>>(b, a, eb);
    } else {
      tdb::mat_bc_kernel<<<blocks < 1184 ? blocks : 1184, 128, 0, s>>>(b);
    }
    MCU(cudaGetLastError());
    if (p->peer && out && !merged) {       // a schedule without the inline exchange: the separate exchange kernel
      const int rc = tdb200_peer_allreduce(p->peer, out, 2 + p->n_slots, s);
      if (rc != TDB200_OK) return rc;
    }
  } else {
    tdb::mat_finalize_kernel<<<1, 256, 0, s>>>(p->d_part_loss, n_ctas, p->desc.n_eq,
                                              (double)p->desc.n0 * (double)p->desc.n1, p->d_bc_sum, p->n_bc_slots,
                                              p->d_slot_lambda, p->d_slot_len, out, p->d_ticket + 1);
    MCU(cudaGetLastError());
  }
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
int32_t tdb200_mat_plan_launches_per_call(const tdb200_mat_plan* p) {
  if (!p) return 0;
  if (p->march && p->edge_bc && p->args.lin1 && getenv("TDB200_MAT_FUSED") && (p->args.edge_x + p->cx_hx) % 4 == 0 &&
      tdb::mat_march_fused_ctas(p->args, p->n_sms) <= 2 * p->n_sms) return 1;
  return p->args.lin1 && p->march && !p->edge_bc ? 3 : 2;
}

int tdb200_mat_plan_set_peer(tdb200_mat_plan* p, tdb200_peer* peer, int32_t* inline_out) {
  if (!p) return mat_invalid("null plan");
  // the loss exchange runs inside the boundary kernel only on the two-launch schedule (mat_march + mat_edge_bc)
  const bool ok = p->march && p->edge_bc && p->n_bc_rows > 0;
  p->peer = ok ? peer : nullptr;
  if (inline_out) *inline_out = ok && peer ? 1 : 0;
  return TDB200_OK;
}

int tdb200_mat_plan_set_row_window(tdb200_mat_plan* p, int32_t row_lo, int32_t row_hi) {
  if (!p) return mat_invalid("null plan");
  if (row_lo < 0 || row_hi > p->desc.n0 || row_lo >= row_hi) return mat_invalid("row window out of range");
  p->args.row_lo = row_lo;
  p->args.row_hi = row_hi;
  return TDB200_OK;
}

int tdb200_mat_plan_set_timing(tdb200_mat_plan* p, int32_t on) {
  if (!p) return mat_invalid("null plan");
  MCU(cudaSetDevice(p->device));
  if (on && !p->ev0) { MCU(cudaEventCreate(&p->ev0)); MCU(cudaEventCreate(&p->ev1)); }
  p->timing = on != 0;
  return TDB200_OK;
}

int tdb200_mat_time_stencil(tdb200_mat_plan* p, const float* u, float* grad, int32_t iters, float* ms_out, void* stream) {
  if (!p || !u || !grad || !ms_out || iters < 1) return mat_invalid("bad argument");
  if (!p->bcs_set) return mat_invalid("tdb200_mat_plan_set_bcs was not called");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MCU(cudaSetDevice(p->device));
  if (!p->ev0) { MCU(cudaEventCreate(&p->ev0)); MCU(cudaEventCreate(&p->ev1)); }
  tdb::MatArgs a = p->args;
  a.u = u; a.grad = grad; a.op_out = nullptr; a.tile_ctr = p->d_ticket + 1;
  int n_ctas = 0;
  MCU(cudaEventRecord(p->ev0, s));
  for (int i = 0; i < iters; ++i) {
    const int rc = mat_stencil(p, a, u, grad, nullptr, &n_ctas, nullptr, true, s);
    if (rc != TDB200_OK) return rc;
    if (p->tma && !p->march) MCU(cudaMemsetAsync(p->d_ticket + 1, 0, sizeof(unsigned int), s));   // re-arm the tile counter
  }
  MCU(cudaEventRecord(p->ev1, s));
  MCU(cudaEventSynchronize(p->ev1));
  float ms = 0.f;
  MCU(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
  *ms_out = ms / (float)iters;
  return TDB200_OK;
}

int tdb200_mat_plan_stencil_ms(tdb200_mat_plan* p, float* ms_out) {
  if (!p || !ms_out) return mat_invalid("null argument");
  if (!p->ev0) return mat_invalid("tdb200_mat_plan_set_timing was not called");
  MCU(cudaEventSynchronize(p->ev1));
  MCU(cudaEventElapsedTime(ms_out, p->ev0, p->ev1));
  return TDB200_OK;
}

int32_t tdb200_mat_plan_kernel_kind(const tdb200_mat_plan* p) {
  if (!p || !p->args.lin1) return 0;
  return p->cross ? (p->march ? 4 : p->tma ? 3 : 2) : 1;
}

void tdb200_mat_plan_destroy(tdb200_mat_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaFree(p->d_band); cudaFree(p->d_terms); cudaFree(p->d_factors); cudaFree(p->d_bcs); cudaFree(p->d_bc_row_begin);
  cudaFree(p->d_slot_scale); cudaFree(p->d_slot_lambda); cudaFree(p->d_slot_len); cudaFree(p->d_part_loss);
  cudaFree(p->d_bc_sum); cudaFree(p->d_ticket); cudaFree(p->d_edge_seed); cudaFree(p->d_csr_off); cudaFree(p->d_csr_ent);
  if (p->ev0) { cudaEventDestroy(p->ev0); cudaEventDestroy(p->ev1); }
  delete p;
}

}  // extern "C"
