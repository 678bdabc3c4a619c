// Flat-arena optimiser pass: gradient norm -> clip -> Adam -> EMA in two launches per network per step.
// Replaces, for the step body of ddgan.py:484-485, 507-508, 517-518, the per-tensor kernels of
// torch.nn.utils.clip_grad_norm_, torch.optim.Adam and ema.py:45-55.
// Pure HBM streaming: 4 reads + 3 writes of fp32 per parameter (+ 2 for the EMA) = 28-36 B/param.
// All scalars that change from step to step (step count, learning rate) live in device memory so that the pass can be
// captured in a CUDA graph.
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

// sum of squares of g[0..n) accumulated into *out (double); *out must be zero before the launch
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n, double* __restrict__ out) {
  double acc = 0.0;
  const long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(g4 + i);
    acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  for (long i = (n4 << 2) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) acc += (double)g[i] * g[i];
  __shared__ double red[8];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// state[0] = step count (float, incremented here by block 0 thread 0 semantics are handled by a separate tiny kernel),
// state[1] = learning rate.
__global__ void adam_tick_kernel(float* state) { state[0] += 1.0f; }

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* ema, float clip, float lr, float b1, float b2,
                                         float eps, float wd, float bc1, float bc2_sqrt, float ema_decay) {
  g *= clip;
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = b1 * m + (1.f - b1) * g;               // torch: exp_avg.lerp_(grad, 1 - beta1)
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - (lr / bc1) * (m / denom);
  if (ema) *ema = ema_decay * (*ema) + (1.f - ema_decay) * p;
}

__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                      float* __restrict__ v, float* __restrict__ ema, long n,
                                                      const float* __restrict__ state, const double* __restrict__ normsq, float max_norm,
                                                      float b1, float b2, float eps, float wd, float ema_decay, float grad_scale) {
  const float step = state[0], lr = state[1];
  // grad_scale (1/world after a sum all-reduce) is applied to the gradient before everything else: the norm of the scaled
  // gradient is grad_scale * sqrt(normsq), and the clip factor multiplies on top of it
  float clip = grad_scale;
  if (normsq != nullptr && max_norm > 0.f) {
    const float total = (float)sqrt(*normsq) * grad_scale;
    clip = grad_scale * fminf(max_norm / (total + 1e-6f), 1.f);   // torch.nn.utils.clip_grad_norm_
  }
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const long n4 = n >> 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 G = ldg_stream(reinterpret_cast<const float4*>(g) + i);
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    float4 E = make_float4(0, 0, 0, 0);
    if (ema) E = reinterpret_cast<float4*>(ema)[i];
    adam_one(P.x, G.x, M.x, V.x, ema ? &E.x : nullptr, clip, lr, b1, b2, eps, wd, bc1, bc2_sqrt, ema_decay);
    adam_one(P.y, G.y, M.y, V.y, ema ? &E.y : nullptr, clip, lr, b1, b2, eps, wd, bc1, bc2_sqrt, ema_decay);
    adam_one(P.z, G.z, M.z, V.z, ema ? &E.z : nullptr, clip, lr, b1, b2, eps, wd, bc1, bc2_sqrt, ema_decay);
    adam_one(P.w, G.w, M.w, V.w, ema ? &E.w : nullptr, clip, lr, b1, b2, eps, wd, bc1, bc2_sqrt, ema_decay);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    if (ema) reinterpret_cast<float4*>(ema)[i] = E;
  }
  for (long i = (n4 << 2) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float P = p[i], M = m[i], V = v[i], E = ema ? ema[i] : 0.f;
    adam_one(P, g[i], M, V, ema ? &E : nullptr, clip, lr, b1, b2, eps, wd, bc1, bc2_sqrt, ema_decay);
    p[i] = P; m[i] = M; v[i] = V;
    if (ema) ema[i] = E;
  }
}

}  // namespace ddg

using namespace ddg;

extern "C" int ddg_grad_norm_sq(const float* g, long n, double* out, cudaStream_t stream) {
  if (!g || !out || n < 0) { ddg_set_last_error("grad_norm_sq: bad args"); return DDG_ERR_ARG; }
  if ((((uintptr_t)g) & 15) != 0) { ddg_set_last_error("grad_norm_sq: arena must be 16-byte aligned"); return DDG_ERR_ARG; }
  cudaMemsetAsync(out, 0, sizeof(double), stream);
  long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(int)blocks, 256, 0, stream>>>(g, n, out);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, long n, float* state, const double* normsq,
                                 float max_norm, float beta1, float beta2, float eps, float weight_decay, float ema_decay,
                                 float grad_scale, cudaStream_t stream) {
  if (!p || !g || !m || !v || !state || n < 0) { ddg_set_last_error("adam_ema_step: bad args"); return DDG_ERR_ARG; }
  if (((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v) | ((uintptr_t)(ema ? ema : p))) & 15) != 0) {
    ddg_set_last_error("adam_ema_step: arenas must be 16-byte aligned");
    return DDG_ERR_ARG;
  }
  adam_tick_kernel<<<1, 1, 0, stream>>>(state);
  long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  adam_ema_kernel<<<(int)blocks, 256, 0, stream>>>(p, g, m, v, ema, n, state, normsq, max_norm, beta1, beta2, eps, weight_decay, ema_decay, grad_scale);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
