// Fused optimizer tail of the train step on ONE flat fp32 buffer per role (parameters, gradients, Adam moments):
// global-norm gradient clipping (runners/pytorch_runner_vae.py:322, torch.nn.utils.clip_grad_norm_) chained with
// the Adam update (:324, optimizer built at :219-220).  Two launches instead of ~40 tensors x (norm, scale, Adam):
//   sumsq_partial_kernel : per-block partial sums of g^2 (fixed grid => deterministic)
//   clip_adam_kernel     : every block re-reduces the partials in a fixed order (bit-identical coefficient in all
//                          blocks), scales the gradient, updates m, v and the parameter; the step counter lives on
//                          the device so the whole step can sit in a CUDA graph.
// HBM-bound: reads p, g, m, v and writes p, m, v (and the clipped g) once: 32 B per parameter.
#pragma once
#include "common.cuh"

namespace acvae {

constexpr int kOptBlocks = 592;     // 4 CTAs per SM x 148 SMs
constexpr int kOptThreads = 256;

__device__ __forceinline__ float block_sum_opt(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  t = red[32];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kOptThreads) sumsq_partial_kernel(long long n4, const float4* __restrict__ g,
                                                                   float* __restrict__ partial) {
  __shared__ float red[33];
  float s = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  s = block_sum_opt(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

struct ClipAdamParams {
  long long n4;            // number of float4 words
  float4* p; float4* g; float4* m; float4* v;
  const float* partial; int npartial;
  float max_norm;          // <= 0: no clipping
  float lr, beta1, beta2, eps, weight_decay;
  const float* hyper;      // optional DEVICE vector {max_norm, lr, beta1, beta2, eps, weight_decay}: overrides the by-value
                           // fields, so an LR scheduler can change them between replays of a captured CUDA graph
  int* step;               // device step counter (incremented by block 0 AFTER every block has read it: see below)
  float* total_norm;       // device scalar out (may be NULL)
  int write_grad;          // store the clipped gradient back (clip_grad_norm_ semantics)
};

__global__ void __launch_bounds__(kOptThreads) clip_adam_kernel(const __grid_constant__ ClipAdamParams a) {
  __shared__ float red[33];
  float s = 0.0f;
  for (int i = threadIdx.x; i < a.npartial; i += blockDim.x) s += a.partial[i];
  s = block_sum_opt(s, red);
  const float norm = sqrtf(s);
  float coef = 1.0f;
  const float max_norm = a.hyper ? a.hyper[0] : a.max_norm;
  if (max_norm > 0.0f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);           // torch: clamp(max_norm / (total + 1e-6), max=1)
  // the counter holds the number of COMPLETED steps; this launch performs step t = counter + 1.  It is advanced by
  // a separate one-thread kernel after this one (no block of this grid may see the new value).
  const int t = a.step[0] + 1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.total_norm) a.total_norm[0] = norm;
  const float lr = a.hyper ? a.hyper[1] : a.lr;
  const float b1 = a.hyper ? a.hyper[2] : a.beta1, b2 = a.hyper ? a.hyper[3] : a.beta2;
  const float eps = a.hyper ? a.hyper[4] : a.eps, wd = a.hyper ? a.hyper[5] : a.weight_decay;
  const float bc1 = 1.0f - powf(b1, (float)t);
  const float bc2 = 1.0f - powf(b2, (float)t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  auto upd = [&](float& p, float& g, float& m, float& v) {
    g *= coef;
    float gg = g;
    if (wd != 0.0f) gg = fmaf(wd, p, gg);                      // Adam's L2 form (torch.optim.Adam, weight_decay)
    m = fmaf(1.0f - b1, gg - m, m);                            // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(1.0f - b2, gg * gg, b2 * v);                      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p -= step_size * (m / denom);
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = a.p[i], g = a.g[i], m = a.m[i], v = a.v[i];
    upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
    if (a.write_grad) a.g[i] = g;
  }
}

__global__ void step_advance_kernel(int* step) { step[0] += 1; }


// ---- loss composition of the runner (runners/pytorch_runner_vae.py:315-320) as one node ----------------------
// terms = {loss, ce, kl, mse}: mse = mean((a - b)^2) over n elements (nn.MSELoss, :318), loss = ce + kl_w*kl + alpha*mse.
// One block, fixed summation order (deterministic).  a == NULL: no global-constraint term.
__global__ void __launch_bounds__(1024) loss_combine_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ ce, const float* __restrict__ kl,
                                                            float kl_w, float alpha, float* __restrict__ terms) {
  __shared__ float red[33];
  float s = 0.0f;
  if (a)
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { const float d = a[i] - b[i]; s = fmaf(d, d, s); }
  s = block_sum_opt(s, red);
  if (threadIdx.x == 0) {
    const float mse = a ? s / (float)n : 0.0f;
    terms[1] = ce[0]; terms[2] = kl[0]; terms[3] = mse;
    terms[0] = ce[0] + kl_w * kl[0] + alpha * mse;
  }
}
// backward: d a = g*alpha*2(a-b)/n, d b = -d a; scal = {g (for the CE rows), g*kl_w (for the KL)}
__global__ void loss_combine_bwd_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b,
                                        const float* __restrict__ g, float kl_w, float alpha, float* __restrict__ da,
                                        float* __restrict__ db, float* __restrict__ scal) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float gg = g[0];
  if (i == 0) { scal[0] = gg; scal[1] = gg * kl_w; }
  if (a && i < n) {
    const float v = gg * alpha * 2.0f * (a[i] - b[i]) / (float)n;
    da[i] = v; db[i] = -v;
  }
}

}  // namespace acvae
