// PNHWC helper kernels of the training path: GroupNorm/AdaGN application (+ activation) and its backward, and the
// per-(sample, channel) statistics with their backward.  Together with the tiny per-(n,c) algebra done on [N,C] tensors they
// are the GroupNorm forward/backward of layerspp.py:46-63 decomposed so that normalisation never needs its own pass over the
// activations in the fused inference plan, and so that autograd composes the exact GroupNorm gradient in training.
// All are HBM-bound: one read (+ one write) of the activation per kernel, 128-bit accesses along the channel axis.
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

// The thread that writes interior pixel (h, w) of a fresh PNHWC buffer also clears the frame pixels next to it (4 channels at c0):
// every frame pixel has exactly one such neighbour, so the one-pixel border ends up zero without a separate launch.
__device__ __forceinline__ void clear_frame(float* ctr, int h, int w, int H, int W, int C) {
  const bool eL = (w == 0), eR = (w == W - 1), eT = (h == 0), eB = (h == H - 1);
  if (!(eL | eR | eT | eB)) return;
  const long rowp = (long)(W + 2) * C;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (eL) *reinterpret_cast<float4*>(ctr - C) = z;
  if (eR) *reinterpret_cast<float4*>(ctr + C) = z;
  if (eT) *reinterpret_cast<float4*>(ctr - rowp) = z;
  if (eB) *reinterpret_cast<float4*>(ctr + rowp) = z;
  if (eT && eL) *reinterpret_cast<float4*>(ctr - rowp - C) = z;
  if (eT && eR) *reinterpret_cast<float4*>(ctr - rowp + C) = z;
  if (eB && eL) *reinterpret_cast<float4*>(ctr + rowp - C) = z;
  if (eB && eR) *reinterpret_cast<float4*>(ctr + rowp + C) = z;
}

__device__ __forceinline__ float act_d(float u, int act) {
  if (act == ACT_SILU) { const float s = 1.f / (1.f + __expf(-u)); return s * (1.f + u * (1.f - s)); }
  if (act == ACT_LEAKY) return u > 0.f ? 1.f : 0.2f;
  return 1.f;
}

__global__ void __launch_bounds__(256) affine_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, float* __restrict__ y, int N, int H,
                                                            int W, int C, int act) {
  const int C4 = C / 4;
  const long total = (long)N * H * W * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long r = i / C4;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const size_t off = ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * C + c4 * 4;
    float4 v = ldg_stream(reinterpret_cast<const float4*>(x + off));
    if (scale) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * C) + c4);
      const float4 t = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * C) + c4);
      v.x = fmaf(v.x, s.x, t.x); v.y = fmaf(v.y, s.y, t.y); v.z = fmaf(v.z, s.z, t.z); v.w = fmaf(v.w, s.w, t.w);
    }
    v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
    stg_stream(reinterpret_cast<float4*>(y + off), v);
    clear_frame(y + off, h, w, H, W, C);
  }
}

// grid (chunks, N); each block covers a contiguous range of interior pixels of sample n for all channels.
// thread -> (pixel lane, channel quad); per-thread partial sums, smem reduce over pixel lanes, one fp64 atomic per channel.
template <int MODE>  // 0: affine_act backward, 1: stats forward
__global__ void __launch_bounds__(256) pnhwc_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          float* __restrict__ dx, double* __restrict__ sums, int H, int W, int C,
                                                          int act, int pix_per_block) {
  extern __shared__ float red[];  // [lanes][C][2]
  const int n = blockIdx.y;
  const int C4 = C / 4;
  const int lanes = blockDim.x / C4;          // pixel lanes (>= 1)
  const int c4 = threadIdx.x % C4;
  const int pl = threadIdx.x / C4;
  const int HW = H * W;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, HW);
  float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  float4 s = make_float4(1, 1, 1, 1), t = make_float4(0, 0, 0, 0);
  if (MODE == 0 && scale) {
    s = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * C) + c4);
    t = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * C) + c4);
  }
  if (pl < lanes) {
    for (int p = p0 + pl; p < p1; p += lanes) {
      const int h = p / W, w = p - h * W;
      const size_t off = ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * C + c4 * 4;
      const float4 v = ldg_stream(reinterpret_cast<const float4*>(x + off));
      if (MODE == 0) {
        const float4 g = ldg_stream(reinterpret_cast<const float4*>(dy + off));
        float4 o;
        float gg;
        gg = g.x * act_d(fmaf(v.x, s.x, t.x), act); a1[0] += gg * v.x; a2[0] += gg; o.x = gg * s.x;
        gg = g.y * act_d(fmaf(v.y, s.y, t.y), act); a1[1] += gg * v.y; a2[1] += gg; o.y = gg * s.y;
        gg = g.z * act_d(fmaf(v.z, s.z, t.z), act); a1[2] += gg * v.z; a2[2] += gg; o.z = gg * s.z;
        gg = g.w * act_d(fmaf(v.w, s.w, t.w), act); a1[3] += gg * v.w; a2[3] += gg; o.w = gg * s.w;
        if (dx) stg_stream(reinterpret_cast<float4*>(dx + off), o);
      } else {
        a1[0] += v.x; a2[0] += v.x * v.x; a1[1] += v.y; a2[1] += v.y * v.y;
        a1[2] += v.z; a2[2] += v.z * v.z; a1[3] += v.w; a2[3] += v.w * v.w;
      }
    }
  }
  if (sums == nullptr) return;
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[((size_t)pl * C + c4 * 4 + j) * 2] = a1[j];
      red[((size_t)pl * C + c4 * 4 + j) * 2 + 1] = a2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += (double)red[(size_t)l * 2 * C + i];
    atomicAdd(sums + (size_t)n * 2 * C + i, acc);
  }
}

__global__ void __launch_bounds__(256) stats_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dx,
                                                       int N, int H, int W, int C) {
  const int C4 = C / 4;
  const long total = (long)N * H * W * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long r = i / C4;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const size_t off = ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * C + c4 * 4;
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(x + off));
    const float* gp = g + ((size_t)n * C + c4 * 4) * 2;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp));      // g1[c], g2[c], g1[c+1], g2[c+1]
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
    float4 o;
    o.x = g0.x + 2.f * v.x * g0.y; o.y = g0.z + 2.f * v.y * g0.w;
    o.z = g1.x + 2.f * v.z * g1.y; o.w = g1.z + 2.f * v.w * g1.w;
    stg_stream(reinterpret_cast<float4*>(dx + off), o);
  }
}

// GroupNorm backward, second pass: dx = dy * act'(scale*x + shift) * scale + g1[n,c] + 2 * x * g2[n,c]
// (the first term is the path through the affine + activation, the other two the path through the statistics)
__global__ void __launch_bounds__(256) gn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       const float* __restrict__ g12, float* __restrict__ dx, int N, int H, int W, int C,
                                                       int act) {
  const int C4 = C / 4;
  const long total = (long)N * H * W * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long r = i / C4;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const size_t off = ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * C + c4 * 4;
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(x + off));
    const float4 g = ldg_stream(reinterpret_cast<const float4*>(dy + off));
    const float4 s = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * C) + c4);
    const float4 t = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * C) + c4);
    const float* gp = g12 + ((size_t)n * C + c4 * 4) * 2;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp));      // g1[c], g2[c], g1[c+1], g2[c+1]
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
    float4 o;
    o.x = g.x * act_d(fmaf(v.x, s.x, t.x), act) * s.x + g0.x + 2.f * v.x * g0.y;
    o.y = g.y * act_d(fmaf(v.y, s.y, t.y), act) * s.y + g0.z + 2.f * v.y * g0.w;
    o.z = g.z * act_d(fmaf(v.z, s.z, t.z), act) * s.z + g1.x + 2.f * v.z * g1.y;
    o.w = g.w * act_d(fmaf(v.w, s.w, t.w), act) * s.w + g1.z + 2.f * v.w * g1.w;
    stg_stream(reinterpret_cast<float4*>(dx + off), o);
    clear_frame(dx + off, h, w, H, W, C);
  }
}

// one thread per (sample, group): from the forward statistics and the first-pass sums {sum gg*x, sum gg} (gg = dy * act'(u)):
// d(gamma), d(beta) per (n, c) and the coefficients g1 = d(sum x), g2 = d(sum x^2) of the second pass.
__global__ void gn_bwd_coeffs_kernel(const double* __restrict__ st, const double* __restrict__ sums, const float* __restrict__ gamma,
                                     int gb_stride, int per_sample, float* __restrict__ g12, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int dgb_stride, int N, int C, int HW, int G, float eps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * G) return;
  const int n = idx / G, g = idx - n * G;
  const int cpg = C / G;
  const size_t base = (size_t)n * C + (size_t)g * cpg;
  double s1 = 0, s2 = 0;
  for (int j = 0; j < cpg; ++j) { s1 += st[(base + j) * 2]; s2 += st[(base + j) * 2 + 1]; }
  const double cnt = (double)cpg * HW;
  const double mean = s1 / cnt;
  double var = s2 / cnt - mean * mean;
  const bool clamped = var < 0;
  if (clamped) var = 0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  double d_rstd = 0, d_mean = 0;
  for (int j = 0; j < cpg; ++j) {
    const int c = g * cpg + j;
    const double ga = gamma ? (double)(per_sample ? gamma[(size_t)n * gb_stride + c] : gamma[c]) : 1.0;
    const double ds = sums[(base + j) * 2], dh = sums[(base + j) * 2 + 1];   // d(scale), d(shift)
    const double t = ds - mean * dh;                 // scale = ga*rstd, shift = be - mean*ga*rstd
    dgamma[(size_t)n * dgb_stride + c] = (float)(rstd * t);
    dbeta[(size_t)n * dgb_stride + c] = (float)dh;
    d_rstd += ga * t;
    d_mean -= rstd * ga * dh;
  }
  const double d_var = clamped ? 0.0 : -0.5 * d_rstd * rstd * rstd * rstd;
  const float d_s1 = (float)(d_mean / cnt - 2.0 * mean * d_var / cnt);
  const float d_s2 = (float)(d_var / cnt);
  for (int j = 0; j < cpg; ++j) { g12[(base + j) * 2] = d_s1; g12[(base + j) * 2 + 1] = d_s2; }
}

}  // namespace ddg

using namespace ddg;

extern "C" int ddg_gn_bwd_dx(const float* x, const float* dy, const float* scale, const float* shift, const float* g12, float* dx, int N,
                             int H, int W, int C, int act, cudaStream_t stream) {
  if (!x || !dy || !scale || !shift || !g12 || !dx || C % 4 != 0) { ddg_set_last_error("gn_bwd_dx: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * H * W * (C / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  gn_bwd_dx_kernel<<<(int)blocks, 256, 0, stream>>>(x, dy, scale, shift, g12, dx, N, H, W, C, act);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_gn_bwd_coeffs(const double* stats, const double* sums, const float* gamma, int gb_stride, int per_sample, float* g12,
                                 float* dgamma, float* dbeta, int dgb_stride, int N, int C, int HW, int G, float eps,
                                 cudaStream_t stream) {
  if (!stats || !sums || !g12 || !dgamma || !dbeta || G <= 0 || C % G != 0) { ddg_set_last_error("gn_bwd_coeffs: bad args"); return DDG_ERR_ARG; }
  const int total = N * G;
  gn_bwd_coeffs_kernel<<<(total + 127) / 128, 128, 0, stream>>>(stats, sums, gamma, gb_stride, per_sample, g12, dgamma, dbeta,
                                                               dgb_stride, N, C, HW, G, eps);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

static int reduce_launch_cfg(int H, int W, int C, int& threads, int& pix_per_block, int& blocks, size_t& smem) {
  const int C4 = C / 4;
  if (C % 4 != 0) return -1;
  if (C4 > 256) return -1;                 // kernels are bounded to 256 threads: one thread per channel quad at least
  threads = (256 / C4) * C4;               // largest multiple of C4 that fits the launch bound
  const int lanes = threads / C4;
  const int HW = H * W;
  // enough blocks to fill the machine: ~4 blocks per SM across the batch, at least 'lanes' pixels per block
  pix_per_block = HW > 64 * lanes ? 16 * lanes : (HW + 3) / 4;
  if (pix_per_block < lanes) pix_per_block = lanes;
  blocks = (HW + pix_per_block - 1) / pix_per_block;
  smem = (size_t)lanes * C * 2 * sizeof(float);
  return 0;
}

extern "C" int ddg_affine_act_fwd(const float* x, const float* scale, const float* shift, float* y, int N, int H, int W, int C, int act,
                                  cudaStream_t stream) {
  if (!x || !y || C % 4 != 0 || ((scale == nullptr) != (shift == nullptr))) { ddg_set_last_error("affine_act_fwd: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * H * W * (C / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  affine_act_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(x, scale, shift, y, N, H, W, C, act);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_affine_act_bwd(const float* x, const float* dy, const float* scale, const float* shift, float* dx, double* sums, int N,
                                  int H, int W, int C, int act, cudaStream_t stream) {
  if (!x || !dy || (!dx && !sums) || ((scale == nullptr) != (shift == nullptr))) { ddg_set_last_error("affine_act_bwd: bad args"); return DDG_ERR_ARG; }
  int threads, ppb, blocks; size_t smem;
  if (reduce_launch_cfg(H, W, C, threads, ppb, blocks, smem)) { ddg_set_last_error("affine_act_bwd: unsupported channel count"); return DDG_ERR_UNSUPPORTED; }
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(pnhwc_reduce_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
  pnhwc_reduce_kernel<0><<<dim3(blocks, N), threads, smem, stream>>>(x, dy, scale, shift, dx, sums, H, W, C, act, ppb);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_stats_fwd(const float* x, double* stats, int N, int H, int W, int C, cudaStream_t stream) {
  if (!x || !stats) { ddg_set_last_error("stats_fwd: bad args"); return DDG_ERR_ARG; }
  int threads, ppb, blocks; size_t smem;
  if (reduce_launch_cfg(H, W, C, threads, ppb, blocks, smem)) { ddg_set_last_error("stats_fwd: unsupported channel count"); return DDG_ERR_UNSUPPORTED; }
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(pnhwc_reduce_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
  pnhwc_reduce_kernel<1><<<dim3(blocks, N), threads, smem, stream>>>(x, nullptr, nullptr, nullptr, nullptr, stats, H, W, C, 0, ppb);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_stats_bwd(const float* x, const float* g, float* dx, int N, int H, int W, int C, cudaStream_t stream) {
  if (!x || !g || !dx || C % 4 != 0) { ddg_set_last_error("stats_bwd: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * H * W * (C / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  stats_bwd_kernel<<<(int)blocks, 256, 0, stream>>>(x, g, dx, N, H, W, C);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
