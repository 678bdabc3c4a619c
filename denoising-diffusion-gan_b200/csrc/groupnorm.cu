// GroupNorm / AdaptiveGroupNorm (+ fused activation) on NCHW, forward and backward, and fused bias + leaky-ReLU.
// Replaces ATen native_group_norm as reached from layerspp.py:46-63,100 and ncsnpp_generator_adagn.py:264, and
// score_sde/op/fused_bias_act_kernel.cu:20-101.
//
// In NCHW the cpg channels of one group are one contiguous run of cpg*HW floats, so a CTA owns one (n, group):
// it streams the run once with 128-bit loads into shared memory (<= 192 KB, i.e. every CIFAR-config group and HQ256
// groups up to 48K elements), reduces with warp shuffles (mean first, then centred sum of squares: exact two-pass
// statistics at one HBM read), and writes y with 128-bit stores.  Algorithmic traffic: 8 B per element (fwd).
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

constexpr int kGnThreads = 512;

template <bool CACHED>
__global__ void __launch_bounds__(kGnThreads) groupnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float* __restrict__ y,
                                                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int C,
                                                                   int HW, int G, float eps, int per_sample, int act) {
  extern __shared__ __align__(16) float cache[];
  __shared__ float red[32];
  const int ng = blockIdx.x;
  const int n = ng / G, g = ng - n * G;
  const int cpg = C / G;
  const long len = (long)cpg * HW;
  const float* xg = x + ((size_t)n * C + (size_t)g * cpg) * HW;
  float* yg = y + ((size_t)n * C + (size_t)g * cpg) * HW;
  const bool vec = ((len & 3) == 0) && ((((uintptr_t)xg | (uintptr_t)yg) & 15) == 0) && ((HW & 3) == 0);
  float s = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xg);
    for (long i = threadIdx.x; i < len / 4; i += blockDim.x) {
      const float4 v = ldg_stream(x4 + i);
      if (CACHED) reinterpret_cast<float4*>(cache)[i] = v;
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (long i = threadIdx.x; i < len; i += blockDim.x) {
      const float v = xg[i];
      if (CACHED) cache[i] = v;
      s += v;
    }
  }
  const float mean = block_sum(s, red) / (float)len;
  float q = 0.f;
  if (vec) {
    for (long i = threadIdx.x; i < len / 4; i += blockDim.x) {
      const float4 v = CACHED ? reinterpret_cast<const float4*>(cache)[i] : __ldg(reinterpret_cast<const float4*>(xg) + i);
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  } else {
    for (long i = threadIdx.x; i < len; i += blockDim.x) {
      const float v = (CACHED ? cache[i] : xg[i]) - mean;
      q += v * v;
    }
  }
  const float var = block_sum(q, red) / (float)len;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[ng] = mean;
    if (rstd_out) rstd_out[ng] = rstd;
  }
  // normalise + affine + activation
  if (vec) {
    const int hw4 = HW / 4;
    for (long i = threadIdx.x; i < len / 4; i += blockDim.x) {
      const int cl = (int)(i / hw4);
      const int c = g * cpg + cl;
      float ga = 1.f, be = 0.f;
      if (gamma) { ga = per_sample ? gamma[(size_t)n * C + c] : gamma[c]; be = per_sample ? beta[(size_t)n * C + c] : beta[c]; }
      const float a = ga * rstd, b = be - mean * a;
      float4 v = CACHED ? reinterpret_cast<const float4*>(cache)[i] : __ldg(reinterpret_cast<const float4*>(xg) + i);
      v.x = apply_act(fmaf(v.x, a, b), act); v.y = apply_act(fmaf(v.y, a, b), act);
      v.z = apply_act(fmaf(v.z, a, b), act); v.w = apply_act(fmaf(v.w, a, b), act);
      stg_stream(reinterpret_cast<float4*>(yg) + i, v);
    }
  } else {
    for (long i = threadIdx.x; i < len; i += blockDim.x) {
      const int c = g * cpg + (int)(i / HW);
      float ga = 1.f, be = 0.f;
      if (gamma) { ga = per_sample ? gamma[(size_t)n * C + c] : gamma[c]; be = per_sample ? beta[(size_t)n * C + c] : beta[c]; }
      const float a = ga * rstd, b = be - mean * a;
      yg[i] = apply_act(fmaf(CACHED ? cache[i] : xg[i], a, b), act);
    }
  }
}

// Backward.  With xhat = (x-mean)*rstd, u = gamma*xhat + beta, y = act(u), dy given:
//   du = dy * act'(u); dgamma[n,c] = sum_hw du*xhat; dbeta[n,c] = sum_hw du
//   dxhat = du*gamma; dx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat*xhat))
__device__ __forceinline__ float act_grad(float u, int act) {
  if (act == ACT_SILU) { const float s = 1.f / (1.f + __expf(-u)); return s * (1.f + u * (1.f - s)); }
  if (act == ACT_LEAKY) return u > 0.f ? 1.f : 0.2f;
  return 1.f;
}

__global__ void __launch_bounds__(kGnThreads) groupnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                   float* __restrict__ dx, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta, int C, int HW, int G, int per_sample,
                                                                   int act) {
  __shared__ float red[32];
  const int ng = blockIdx.x;
  const int n = ng / G, g = ng - n * G;
  const int cpg = C / G;
  const long len = (long)cpg * HW;
  const size_t base = ((size_t)n * C + (size_t)g * cpg) * HW;
  const float mean = mean_in[ng], rstd = rstd_in[ng];
  // pass 1: per-channel sums of du and du*xhat (needed for dgamma/dbeta and for the group means)
  float sum1 = 0.f, sum2 = 0.f;  // sum dxhat, sum dxhat*xhat over the group
  for (int cl = 0; cl < cpg; ++cl) {
    const int c = g * cpg + cl;
    float ga = 1.f, be = 0.f;
    if (gamma) { ga = per_sample ? gamma[(size_t)n * C + c] : gamma[c]; be = per_sample ? beta[(size_t)n * C + c] : beta[c]; }
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float xh = (x[base + (size_t)cl * HW + i] - mean) * rstd;
      const float du = dy[base + (size_t)cl * HW + i] * act_grad(fmaf(ga, xh, be), act);
      a += du * xh;
      b += du;
    }
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0) {
      if (dgamma) dgamma[(size_t)n * C + c] = a;
      if (dbeta) dbeta[(size_t)n * C + c] = b;
    }
    sum1 += b * ga;
    sum2 += a * ga;
  }
  const float m1 = sum1 / (float)len, m2 = sum2 / (float)len;
  for (long i = threadIdx.x; i < len; i += blockDim.x) {
    const int c = g * cpg + (int)(i / HW);
    float ga = 1.f, be = 0.f;
    if (gamma) { ga = per_sample ? gamma[(size_t)n * C + c] : gamma[c]; be = per_sample ? beta[(size_t)n * C + c] : beta[c]; }
    const float xh = (x[base + i] - mean) * rstd;
    const float du = dy[base + i] * act_grad(fmaf(ga, xh, be), act);
    dx[base + i] = rstd * (du * ga - m1 - xh * m2);
  }
}

// fused_bias_act_kernel.cu:20-51, 128-bit vectorised when step_b % 4 == 0
__global__ void __launch_bounds__(256) fused_bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                             const float* __restrict__ ref, float* __restrict__ y, long n, int step_b,
                                                             int size_b, int act, int grad, float alpha, float scale, int vec) {
  const long stride = (long)gridDim.x * blockDim.x;
  if (vec) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n / 4; i += stride) {
      float4 v = ldg_stream(reinterpret_cast<const float4*>(x) + i);
      if (b) { const float bb = __ldg(b + ((i * 4) / step_b) % size_b); v.x += bb; v.y += bb; v.z += bb; v.w += bb; }
      float4 r = make_float4(0, 0, 0, 0);
      if (ref) r = ldg_stream(reinterpret_cast<const float4*>(ref) + i);
      float4 o;
      if (act == 3) {
        if (grad == 0) { o.x = v.x > 0 ? v.x : v.x * alpha; o.y = v.y > 0 ? v.y : v.y * alpha; o.z = v.z > 0 ? v.z : v.z * alpha; o.w = v.w > 0 ? v.w : v.w * alpha; }
        else if (grad == 1) { o.x = r.x > 0 ? v.x : v.x * alpha; o.y = r.y > 0 ? v.y : v.y * alpha; o.z = r.z > 0 ? v.z : v.z * alpha; o.w = r.w > 0 ? v.w : v.w * alpha; }
        else o = make_float4(0, 0, 0, 0);
      } else {
        if (grad == 2) o = make_float4(0, 0, 0, 0); else o = v;
      }
      o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale;
      stg_stream(reinterpret_cast<float4*>(y) + i, o);
    }
  } else {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += stride) {
      float v = x[i];
      if (b) v += b[(i / step_b) % size_b];
      const float r = ref ? ref[i] : 0.f;
      float o;
      if (act == 3) o = grad == 0 ? (v > 0 ? v : v * alpha) : (grad == 1 ? (r > 0 ? v : v * alpha) : 0.f);
      else o = grad == 2 ? 0.f : v;
      y[i] = o * scale;
    }
  }
}

}  // namespace ddg

using namespace ddg;

extern "C" int ddg_groupnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int N, int C,
                                 int HW, int G, float eps, int per_sample, int act, cudaStream_t stream) {
  if (!x || !y || N <= 0 || C <= 0 || HW <= 0 || G <= 0 || C % G != 0 || ((gamma == nullptr) != (beta == nullptr))) {
    ddg_set_last_error("groupnorm_fwd: bad args");
    return DDG_ERR_ARG;
  }
  const size_t bytes = (size_t)(C / G) * HW * sizeof(float);
  if (bytes <= 192 * 1024) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(groupnorm_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024); attr = true; }
    groupnorm_fwd_kernel<true><<<N * G, kGnThreads, bytes, stream>>>(x, gamma, beta, y, mean, rstd, C, HW, G, eps, per_sample, act);
  } else {
    groupnorm_fwd_kernel<false><<<N * G, kGnThreads, 0, stream>>>(x, gamma, beta, y, mean, rstd, C, HW, G, eps, per_sample, act);
  }
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_groupnorm_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* mean,
                                 const float* rstd, float* dx, float* dgamma_nc, float* dbeta_nc, int N, int C, int HW, int G,
                                 int per_sample, int act, cudaStream_t stream) {
  if (!x || !dy || !mean || !rstd || !dx || C % G != 0) { ddg_set_last_error("groupnorm_bwd: bad args"); return DDG_ERR_ARG; }
  groupnorm_bwd_kernel<<<N * G, kGnThreads, 0, stream>>>(x, dy, gamma, beta, mean, rstd, dx, dgamma_nc, dbeta_nc, C, HW, G, per_sample, act);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_fused_bias_act(const float* x, const float* b, const float* ref, float* y, long n, int step_b, int size_b, int act,
                                  int grad, float alpha, float scale, cudaStream_t stream) {
  if (n == 0) return DDG_OK;
  if (!x || !y || n < 0 || (b && (step_b <= 0 || size_b <= 0))) { ddg_set_last_error("fused_bias_act: bad args"); return DDG_ERR_ARG; }
  const int vec = ((n & 3) == 0) && (!b || (step_b % 4 == 0)) && ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)(ref ? ref : x)) & 15) == 0);
  long work = vec ? n / 4 : n;
  long blocks = (work + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  fused_bias_act_kernel<<<(int)blocks, 256, 0, stream>>>(x, b, ref, y, n, step_b, size_b, act, grad, alpha, scale, vec);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
