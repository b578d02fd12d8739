// Fused optimiser step (reference: tedeous/optimizers/optimizer.py:44-61 hands torch.optim.Adam / AdamW /
// SGD the gradient that loss.backward() wrote): ONE launch updates every parameter tensor in place from the flat
// gradient that tdb200_loss_grad / tdb200_mat_loss_grad left in the output vector, so a whole training step
// (pack -> jet kernels -> reduce -> [all-reduce] -> update) is a fixed launch sequence that replays as one CUDA graph.
// Step count and learning rate live in device memory (bias correction needs t; schedulers change lr between replays).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>

#include "tdb200.h"

extern "C" void tdb200_set_error_(const char* msg);

namespace {

struct OptArgs {
  float* p[TDB200_OPT_MAX_TENSORS];
  int64_t begin[TDB200_OPT_MAX_TENSORS + 1];    // tensor i owns flat indices [begin[i], begin[i + 1])
  int n_tensors;
  int kind;                                      // 0 Adam (L2 weight decay), 1 AdamW (decoupled), 2 SGD (momentum)
  const float* grad;                             // flat gradient, same order as the tensors
  float* m;                                      // Adam: first moment / SGD: momentum buffer
  float* v;                                      // Adam: second moment
  const float* hyper;                            // [lr, beta1 | momentum, beta2, eps, weight_decay]
  const int* step;                               // optimiser steps taken so far
};

__global__ void __launch_bounds__(256) fused_optimizer_kernel(const OptArgs a) {
  const int64_t n = a.begin[a.n_tensors];
  const float lr = a.hyper[0], b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], wd = a.hyper[4];
  const int t = a.step[0] + 1;
  float c1 = 1.f, c2s = 1.f;
  if (a.kind != 2) {                             // torch.optim.Adam: step_size = lr / (1 - b1^t), denom = sqrt(v) / sqrt(1 - b2^t) + eps
    c1 = 1.f - powf(b1, (float)t);
    c2s = sqrtf(1.f - powf(b2, (float)t));
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int ti = 0;
    while (i >= a.begin[ti + 1]) ++ti;           // <= 34 tensors: a short scan
    float* q = a.p[ti] + (i - a.begin[ti]);
    float w = *q, g = a.grad[i];
    if (a.kind == 2) {
      g = fmaf(wd, w, g);
      float buf = g;
      if (b1 != 0.f) {
        buf = t == 1 ? g : fmaf(b1, a.m[i], g);
        a.m[i] = buf;
      }
      w -= lr * buf;
    } else {
      if (a.kind == 1) w *= 1.f - lr * wd;
      else g = fmaf(wd, w, g);
      const float m = fmaf(b1, a.m[i], (1.f - b1) * g);
      const float v = fmaf(b2, a.v[i], (1.f - b2) * g * g);
      a.m[i] = m;
      a.v[i] = v;
      w -= (lr / c1) * m / (sqrtf(v) / c2s + eps);
    }
    *q = w;
  }
}

__global__ void optimizer_tick_kernel(int* step) { step[0] += 1; }

}  // namespace

extern "C" int tdb200_optimizer_step(int32_t kind, int32_t n_tensors, float* const* params_dev, const int64_t* sizes,
                                     const float* grad_dev, float* m_dev, float* v_dev, const float* hyper_dev,
                                     int32_t* step_dev, void* stream) {
  if (kind < 0 || kind > 2 || n_tensors < 1 || n_tensors > TDB200_OPT_MAX_TENSORS || !params_dev || !sizes || !grad_dev ||
      !hyper_dev || !step_dev || (kind != 2 && (!m_dev || !v_dev))) {
    tdb200_set_error_("tdb200_optimizer_step: bad argument");
    return TDB200_ERR_INVALID;
  }
  OptArgs a{};
  a.n_tensors = n_tensors;
  a.kind = kind;
  a.begin[0] = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (!params_dev[i] || sizes[i] < 0) { tdb200_set_error_("tdb200_optimizer_step: null parameter tensor"); return TDB200_ERR_INVALID; }
    a.p[i] = params_dev[i];
    a.begin[i + 1] = a.begin[i] + sizes[i];
  }
  a.grad = grad_dev; a.m = m_dev; a.v = v_dev; a.hyper = hyper_dev; a.step = step_dev;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = a.begin[n_tensors];
  int grid = (int)((n + 255) / 256);
  grid = grid < 1 ? 1 : (grid > 592 ? 592 : grid);
  fused_optimizer_kernel<<<grid, 256, 0, s>>>(a);
  optimizer_tick_kernel<<<1, 1, 0, s>>>(step_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { tdb200_set_error_((std::string("tdb200_optimizer_step: ") + cudaGetErrorString(e)).c_str()); return TDB200_ERR_CUDA; }
  return TDB200_OK;
}
