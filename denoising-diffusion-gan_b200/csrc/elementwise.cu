// Small fused kernels of the DDGAN hot path: embeddings, linear layers on [N, K] rows, the Gaussian updates of the
// diffusion process, layout conversion and GroupNorm coefficient preparation.  All HBM-bound or latency-bound.
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

// layers.py:475-486
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ out, int N, int dim, float log_max) {
  pdl_wait();
  pdl_trigger();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * dim) return;
  const int n = i / dim, j = i - n * dim;
  float v = 0.f;
  if (j < 2 * half) {
    const int jj = j < half ? j : j - half;
    // same operation order as the reference: exp(arange * -(log(max)/(half-1))), then t * freq
    const float freq = expf((float)jj * -(log_max / (float)(half - 1)));
    const float arg = (float)t[n] * freq;
    v = j < half ? sinf(arg) : cosf(arg);
  }
  out[i] = v;
}

// y[n][j] = act_out(sum_k act_in(x[n][k]) W[j][k] + b[j]); one warp per output column, rows tiled by 32 in smem.
constexpr int kLinRows = 32;
constexpr int kLinKC = 1024;   // K chunk of the long-K variant
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                     const float* __restrict__ b, float* __restrict__ y, int N, int K, int J,
                                                     int ldx, int ldy, int act_in, int act_out, int pixel_norm, int cpw) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sx[];  // [kLinRows][K]
  const int n0 = blockIdx.y * kLinRows;
  const int rows = min(kLinRows, N - n0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < rows; r += 8) {
    const float* xr = x + (size_t)(n0 + r) * ldx;
    float nrm = 1.f;
    if (pixel_norm) {
      float ss = 0.f;
      for (int k = lane; k < K; k += 32) { const float v = xr[k]; ss += v * v; }
      ss = warp_sum(ss);
      nrm = 1.0f / sqrtf(ss / (float)K + 1e-8f);
    }
    for (int k = lane; k < K; k += 32) sx[r * K + k] = apply_act(xr[k] * nrm, act_in);
  }
  __syncthreads();
  const int jpb = 8 * cpw;  // cpw columns per warp per block (1 for narrow layers: more CTAs)
  for (int jj = 0; jj < cpw; ++jj) {
    const int j = blockIdx.x * jpb + warp * cpw + jj;
    if (j >= J) break;
    const float* wr = W + (size_t)j * K;
    const float bj = b ? b[j] : 0.f;
    for (int r = 0; r < rows; ++r) {
      float acc = 0.f;
      for (int k = lane; k < K; k += 32) acc = fmaf(sx[r * K + k], __ldg(wr + k), acc);
      acc = warp_sum(acc);
      if (lane == 0) y[(size_t)(n0 + r) * ldy + j] = apply_act(acc + bj, act_out);
    }
  }
}

// long-K variant (K > ~1500: the input-gradient GEMM of the batched style projection, K = 2 * sum(C) ~ 31 K):
// K is streamed through shared memory in chunks; each lane keeps a partial sum per row in registers.
__global__ void __launch_bounds__(256) linear_longk_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ b, float* __restrict__ y, int N, int K, int J,
                                                           int ldx, int ldy, int act_in, int act_out) {
  extern __shared__ float sx[];  // [kLinRows][kLinKC]
  const int n0 = blockIdx.y * kLinRows;
  const int rows = min(kLinRows, N - n0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  float acc[kLinRows];
#pragma unroll
  for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kLinKC) {
    const int kc = min(kLinKC, K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * kc; i += blockDim.x) {
      const int r = i / kc, k = i - r * kc;
      sx[r * kLinKC + k] = apply_act(x[(size_t)(n0 + r) * ldx + k0 + k], act_in);
    }
    __syncthreads();
    if (j < J) {
      const float* wr = W + (size_t)j * K + k0;
      for (int k = lane; k < kc; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) acc[r] = fmaf(sx[r * kLinKC + k], wv, acc[r]);
      }
    }
  }
  if (j < J) {
    const float bj = b ? b[j] : 0.f;
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) {
      const float v = warp_sum(acc[r]);
      if (lane == 0 && r < rows) y[(size_t)(n0 + r) * ldy + j] = apply_act(v + bj, act_out);
    }
  }
}

// ddgan.py:110-126
__global__ void q_sample_pairs_kernel(const float4* __restrict__ x0, const float4* __restrict__ n0, const float4* __restrict__ n1,
                                      const int64_t* __restrict__ t, const float* __restrict__ a_cum, const float* __restrict__ s_cum,
                                      const float* __restrict__ a_s, const float* __restrict__ sig, float4* __restrict__ xt,
                                      float4* __restrict__ xtp1, long per4, long total4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per4);
    const int tt = (int)t[n];
    const float ac = a_cum[tt], sc = s_cum[tt], a1 = a_s[tt + 1], s1 = sig[tt + 1];
    const float4 x = ldg_stream(x0 + i), e0 = ldg_stream(n0 + i), e1 = ldg_stream(n1 + i);
    float4 a, b;
    a.x = ac * x.x + sc * e0.x; a.y = ac * x.y + sc * e0.y; a.z = ac * x.z + sc * e0.z; a.w = ac * x.w + sc * e0.w;
    b.x = a1 * a.x + s1 * e1.x; b.y = a1 * a.y + s1 * e1.y; b.z = a1 * a.z + s1 * e1.z; b.w = a1 * a.w + s1 * e1.w;
    stg_stream(xt + i, a);
    stg_stream(xtp1 + i, b);
  }
}

// ddgan.py:152-169: mean = c1[t] x0 + c2[t] x_t ; out = mean + (t != 0) * exp(0.5 logvar[t]) * noise
__global__ void sample_posterior_kernel(const float4* __restrict__ x0, const float4* __restrict__ xt, const float4* __restrict__ nz,
                                        const int64_t* __restrict__ t, const float* __restrict__ c1, const float* __restrict__ c2,
                                        const float* __restrict__ logvar, float4* __restrict__ out, long per4, long total4) {
  pdl_wait();
  pdl_trigger();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per4);
    const int tt = (int)t[n];
    const float k1 = c1[tt], k2 = c2[tt];
    const float sd = (tt != 0 ? 1.f : 0.f) * expf(0.5f * logvar[tt]);
    const float4 a = ldg_stream(x0 + i), b = ldg_stream(xt + i), e = ldg_stream(nz + i);
    float4 o;
    o.x = (k1 * a.x + k2 * b.x) + sd * e.x;
    o.y = (k1 * a.y + k2 * b.y) + sd * e.y;
    o.z = (k1 * a.z + k2 * b.z) + sd * e.z;
    o.w = (k1 * a.w + k2 * b.w) + sd * e.w;
    stg_stream(out + i, o);
  }
}

// scalar-tail variants (per_sample not a multiple of 4)
__global__ void q_sample_pairs_scalar(const float* x0, const float* n0, const float* n1, const int64_t* t, const float* a_cum,
                                      const float* s_cum, const float* a_s, const float* sig, float* xt, float* xtp1, long per,
                                      long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int tt = (int)t[i / per];
    const float a = a_cum[tt] * x0[i] + s_cum[tt] * n0[i];
    xt[i] = a;
    xtp1[i] = a_s[tt + 1] * a + sig[tt + 1] * n1[i];
  }
}
__global__ void sample_posterior_scalar(const float* x0, const float* xt, const float* nz, const int64_t* t, const float* c1,
                                        const float* c2, const float* logvar, float* out, long per, long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int tt = (int)t[i / per];
    const float sd = (tt != 0 ? 1.f : 0.f) * expf(0.5f * logvar[tt]);
    out[i] = (c1[tt] * x0[i] + c2[tt] * xt[i]) + sd * nz[i];
  }
}

// NCHW (a ++ b along C) -> PNHWC interior, channels >= Ca+Cb zero.  One thread per (n, h, w); writes Cpad floats.
__global__ void nchw_to_pnhwc_kernel(const float* __restrict__ a, int Ca, const float* __restrict__ b, int Cb, float* __restrict__ out,
                                     int N, int H, int W, int Cpad, float scale, float shift) {
  pdl_wait();
  pdl_trigger();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)N * H * W * (Cpad / 4);
  if (i >= total) return;
  const int c4 = (int)(i % (Cpad / 4));
  long r = i / (Cpad / 4);
  const int w = (int)(r % W); r /= W;
  const int h = (int)(r % H);
  const int n = (int)(r / H);
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c4 * 4 + j;
    float x = 0.f;
    if (c < Ca) x = a[((size_t)(n * Ca + c) * H + h) * W + w] * scale + shift;
    else if (c < Ca + Cb) x = b[((size_t)(n * Cb + (c - Ca)) * H + h) * W + w] * scale + shift;
    v[j] = x;
  }
  float4* o = reinterpret_cast<float4*>(out + ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * Cpad) + c4;
  *o = make_float4(v[0], v[1], v[2], v[3]);
}

// PNHWC / NHWC -> NCHW through a 32x32 smem transpose (coalesced on both sides)
__global__ void pnhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C, int Cpitch,
                                     int padded) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int HW = H * W;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) {
      const int h = p / W, w = p - h * W;
      const size_t pix = padded ? ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) : ((size_t)n * HW + p);
      v = x[pix * Cpitch + c];
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) out[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

// GroupNorm coefficients from per-(n,c) {sum, sumsq}: one block per sample, one thread per channel.
__global__ void gn_prepare_kernel(const double* __restrict__ sa, int Ca, const double* __restrict__ sb, int Cb,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int gb_stride, int per_sample,
                                  float* __restrict__ scale, float* __restrict__ shift, int HW, int G, float eps) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double sred[];  // [2*C]
  const int n = blockIdx.x;
  const int C = Ca + Cb;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double* s = c < Ca ? sa + ((size_t)n * Ca + c) * 2 : sb + ((size_t)n * Cb + (c - Ca)) * 2;
    sred[2 * c] = s[0];
    sred[2 * c + 1] = s[1];
  }
  __syncthreads();
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    double s1 = 0, s2 = 0;
    for (int j = 0; j < cpg; ++j) { s1 += sred[2 * (g * cpg + j)]; s2 += sred[2 * (g * cpg + j) + 1]; }
    const double cnt = (double)cpg * HW;
    const double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float ga = 1.f, be = 0.f;
    if (gamma) {
      ga = per_sample ? gamma[(size_t)n * gb_stride + c] : gamma[c];
      be = per_sample ? beta[(size_t)n * gb_stride + c] : beta[c];
    }
    const float sc = ga * rstd;
    scale[(size_t)n * C + c] = sc;
    shift[(size_t)n * C + c] = be - (float)mean * sc;
  }
}

// out[n][c] = sum over interior pixels of act(x)
__global__ void spatial_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W, int C, int act) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int h = 0; h < H; ++h)
    for (int w = 0; w < W; ++w) acc += apply_act(x[((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * C + c], act);
  out[(size_t)n * C + c] = acc;
}

// discriminator.py:150-158.  group = min(N, 4); M = N / group; sample i belongs to stat-set (i % M);
// std over the `group` members, mean over (C, H, W) -> one scalar per stat-set, broadcast to channel 0 of out.
__global__ void minibatch_stddev_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C, int Cpad,
                                        int group) {
  const int M = N / group;
  const int m = blockIdx.x;
  const long per = (long)H * W * C;
  double acc = 0.0;
  for (long i = threadIdx.x; i < per; i += blockDim.x) {
    const int c = (int)(i % C);
    long r = i / C;
    const int w = (int)(r % W);
    const int h = (int)(r / W);
    float vals[4];
    float mean = 0.f;
    for (int g = 0; g < group; ++g) {
      const int n = g * M + m;
      vals[g] = x[((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * C + c];
      mean += vals[g];
    }
    mean /= (float)group;
    float var = 0.f;
    for (int g = 0; g < group; ++g) { const float d = vals[g] - mean; var += d * d; }
    var /= (float)group;
    acc += (double)sqrtf(var + 1e-8f);
  }
  __shared__ double red[32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) red[0] = v / (double)per;
  }
  __syncthreads();
  const float s = (float)red[0];
  for (int i = threadIdx.x; i < group * H * W; i += blockDim.x) {
    const int g = i / (H * W);
    const int r = i - g * H * W;
    const int h = r / W, w = r - h * W;
    const int n = g * M + m;
    float* o = out + ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * Cpad;
    o[0] = s;
  }
}

// grad_bias reduction: out[c] = sum_{n, inner} g[n][c][inner]
__global__ void channel_sum_kernel(const float* __restrict__ g, float* __restrict__ out, int N, int C, int inner) {
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int n = 0; n < N; ++n) {
    const float* p = g + ((size_t)n * C + c) * inner;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) acc += (double)p[i];
  }
  __shared__ double red[32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) out[c] = (float)v;
  }
}

}  // namespace ddg

using namespace ddg;

static inline int grid_for(long total, int threads, int cap = 148 * 32) {
  long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

extern "C" int ddg_timestep_embedding(const int64_t* t, float* out, int N, int dim, float max_positions, cudaStream_t stream) {
  if (!t || !out || N <= 0 || dim < 4) { ddg_set_last_error("timestep_embedding: bad args"); return DDG_ERR_ARG; }
  const int total = N * dim;
  launch_pdl(timestep_embedding_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, t, out, N, dim, logf(max_positions));
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_linear(const float* x, const float* W, const float* b, float* y, int N, int K, int J, int ldx, int ldy,
                          int act_in, int act_out, int pixel_norm, cudaStream_t stream) {
  if (!x || !W || !y || N <= 0 || K <= 0 || J <= 0) { ddg_set_last_error("linear: bad args"); return DDG_ERR_ARG; }
  const size_t smem = (size_t)kLinRows * K * sizeof(float);
  if (smem > 200 * 1024) {
    if (pixel_norm) { ddg_set_last_error("linear: pixel_norm needs K <= 1600"); return DDG_ERR_UNSUPPORTED; }
    static bool attr2 = false;
    if (!attr2) { cudaFuncSetAttribute(linear_longk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr2 = true; }
    dim3 grid2((J + 7) / 8, (N + kLinRows - 1) / kLinRows);
    linear_longk_kernel<<<grid2, 256, (size_t)kLinRows * kLinKC * sizeof(float), stream>>>(x, W, b, y, N, K, J, ldx, ldy, act_in, act_out);
    DDG_CHECK_LAUNCH();
    return DDG_OK;
  }
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  const int cpw = J <= 2048 ? 1 : 4;
  dim3 grid((J + 8 * cpw - 1) / (8 * cpw), (N + kLinRows - 1) / kLinRows);
  launch_pdl(linear_kernel, dim3(grid), dim3(256), smem, stream, x, W, b, y, N, K, J, ldx, ldy, act_in, act_out, pixel_norm, cpw);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

namespace ddg {
// One CTA = kMlpRows rows through every layer; activations ping-pong in shared memory.  Thread j owns output j of the current
// layer and streams row j of W straight from L2 with 128-bit loads (8 in flight), multiplying it into the CTA's rows: one pass
// over each weight matrix per CTA, no intra-layer barriers.  The chain is latency-bound, so rows are spread thin (2 per CTA,
// 32 CTAs at batch 64: 48 us for the 5-layer z network, 15 us per 256 x 256 layer; 8 rows per CTA: 63 us; a shared-memory
// weight-tile version with two barriers per 64 x 64 tile: 214 us; warp-per-output with coalesced loads + shuffles: 133 us).
constexpr int kMlpRows = 2;
constexpr int kMlpMaxDim = 1024;
__global__ void __launch_bounds__(256) mlp_rows_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int N,
                                                       const ddg_mlp_desc d) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float msm[];       // act[2][kMlpRows][kMlpMaxDim]
  float* act0 = msm;
  float* act1 = msm + kMlpRows * kMlpMaxDim;
  const int n0 = blockIdx.x * kMlpRows;
  const int rows = min(kMlpRows, N - n0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int K0 = d.dims[0];
    if (warp < kMlpRows) {
      float nrm = 1.f;
      if (warp < rows) {
        const float* xr = x + (size_t)(n0 + warp) * ldx;
        if (d.pixel_norm) {
          float ss = 0.f;
          for (int k = lane; k < K0; k += 32) { const float v = xr[k]; ss += v * v; }
          ss = warp_sum(ss);
          nrm = 1.0f / sqrtf(ss / (float)K0 + 1e-8f);
        }
        for (int k = lane; k < K0; k += 32) act0[warp * kMlpMaxDim + k] = xr[k] * nrm;
      } else {
        for (int k = lane; k < K0; k += 32) act0[warp * kMlpMaxDim + k] = 0.f;
      }
    }
  }
  __syncthreads();
  float* cur = act0;
  float* nxt = act1;
#pragma unroll 1
  for (int l = 0; l < DDG_MLP_MAX_LAYERS; ++l) {
    if (l >= d.nlayers) break;
    const int K = d.dims[l], J = d.dims[l + 1];
    const float* __restrict__ W = d.W[l];
    const float* __restrict__ b = d.b[l];
    const bool last = (l == d.nlayers - 1);
    const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    // (a warp-per-output mapping with coalesced weight loads and a shuffle reduction measured 1.7x slower: what matters at 8 CTAs
    // is the number of independent loads in flight per thread, not coalescing -- the weights are L2-resident)
    for (int j = threadIdx.x; j < J; j += 256) {
      float acc[kMlpRows];
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.f;
      if (vec) {
        const float4* wr = reinterpret_cast<const float4*>(W + (size_t)j * K);
        const int K4 = K / 4;
        int k4 = 0;
        for (; k4 + 8 <= K4; k4 += 8) {
          float4 w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = __ldg(wr + k4 + u);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) {
              const float4 a = *reinterpret_cast<const float4*>(cur + r * kMlpMaxDim + (k4 + u) * 4);
              acc[r] = fmaf(a.x, w[u].x, fmaf(a.y, w[u].y, fmaf(a.z, w[u].z, fmaf(a.w, w[u].w, acc[r]))));
            }
          }
        }
        for (; k4 < K4; ++k4) {
          const float4 w = __ldg(wr + k4);
#pragma unroll
          for (int r = 0; r < kMlpRows; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(cur + r * kMlpMaxDim + k4 * 4);
            acc[r] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[r]))));
          }
        }
      } else {
        for (int k = 0; k < K; ++k) {
          const float w = __ldg(W + (size_t)j * K + k);
#pragma unroll
          for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(cur[r * kMlpMaxDim + k], w, acc[r]);
        }
      }
      const float bj = b ? __ldg(b + j) : 0.f;
#pragma unroll
      for (int r = 0; r < kMlpRows; ++r) {
        const float v = acc[r] + bj;
        if (last) { if (r < rows) y[(size_t)(n0 + r) * ldy + j] = v; }
        else nxt[r * kMlpMaxDim + j] = apply_act(v, d.act);
      }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}
}  // namespace ddg

extern "C" int ddg_mlp_rows(const float* x, int ldx, float* y, int ldy, int N, const ddg_mlp_desc* desc, cudaStream_t stream) {
  if (!x || !y || !desc || N <= 0 || desc->nlayers < 1 || desc->nlayers > DDG_MLP_MAX_LAYERS) { ddg_set_last_error("mlp_rows: bad args"); return DDG_ERR_ARG; }
  for (int i = 0; i <= desc->nlayers; ++i)
    if (desc->dims[i] < 1 || desc->dims[i] > ddg::kMlpMaxDim) { ddg_set_last_error("mlp_rows: layer width out of range (1..1024)"); return DDG_ERR_UNSUPPORTED; }
  for (int i = 0; i < desc->nlayers; ++i)
    if (!desc->W[i]) { ddg_set_last_error("mlp_rows: null weight"); return DDG_ERR_ARG; }
  const size_t smem = (size_t)(2 * ddg::kMlpRows * ddg::kMlpMaxDim) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(ddg::mlp_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  ddg::launch_pdl(ddg::mlp_rows_kernel, dim3((N + ddg::kMlpRows - 1) / ddg::kMlpRows), dim3(256), smem, stream, x, ldx, y, ldy, N, *desc);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_q_sample_pairs(const float* x0, const float* noise_xt, const float* noise_xtp1, const int64_t* t,
                                  const float* a_s_cum, const float* sigmas_cum, const float* a_s, const float* sigmas, float* x_t,
                                  float* x_tp1, int N, long per_sample, cudaStream_t stream) {
  if (!x0 || !noise_xt || !noise_xtp1 || !t || !x_t || !x_tp1 || N <= 0 || per_sample <= 0) { ddg_set_last_error("q_sample_pairs: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * per_sample;
  const bool vec = (per_sample % 4 == 0) && ((((uintptr_t)x0 | (uintptr_t)noise_xt | (uintptr_t)noise_xtp1 | (uintptr_t)x_t | (uintptr_t)x_tp1) & 15) == 0);
  if (vec)
    q_sample_pairs_kernel<<<grid_for(total / 4, 256), 256, 0, stream>>>((const float4*)x0, (const float4*)noise_xt, (const float4*)noise_xtp1, t, a_s_cum,
                                                                       sigmas_cum, a_s, sigmas, (float4*)x_t, (float4*)x_tp1, per_sample / 4, total / 4);
  else
    q_sample_pairs_scalar<<<grid_for(total, 256), 256, 0, stream>>>(x0, noise_xt, noise_xtp1, t, a_s_cum, sigmas_cum, a_s, sigmas, x_t, x_tp1, per_sample, total);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_sample_posterior(const float* x0, const float* x_t, const float* noise, const int64_t* t, const float* coef1,
                                    const float* coef2, const float* logvar, float* out, int N, long per_sample, cudaStream_t stream) {
  if (!x0 || !x_t || !noise || !t || !out || N <= 0 || per_sample <= 0) { ddg_set_last_error("sample_posterior: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * per_sample;
  const bool vec = (per_sample % 4 == 0) && ((((uintptr_t)x0 | (uintptr_t)x_t | (uintptr_t)noise | (uintptr_t)out) & 15) == 0);
  if (vec)
    launch_pdl(sample_posterior_kernel, dim3(grid_for(total / 4, 256)), dim3(256), 0, stream, (const float4*)x0, (const float4*)x_t, (const float4*)noise, t,
               coef1, coef2, logvar, (float4*)out, per_sample / 4, total / 4);
  else
    sample_posterior_scalar<<<grid_for(total, 256), 256, 0, stream>>>(x0, x_t, noise, t, coef1, coef2, logvar, out, per_sample, total);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_nchw_to_pnhwc(const float* a, int Ca, const float* b, int Cb, float* out, int N, int H, int W, int Cpad, float scale,
                                 float shift, cudaStream_t stream) {
  if (!a || !out || Cpad % 4 != 0 || Ca + Cb > Cpad) { ddg_set_last_error("nchw_to_pnhwc: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * H * W * (Cpad / 4);
  launch_pdl(nchw_to_pnhwc_kernel, dim3((int)((total + 255) / 256)), dim3(256), 0, stream, a, Ca, b, b ? Cb : 0, out, N, H, W, Cpad, scale, shift);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_pnhwc_to_nchw(const float* x, float* out, int N, int H, int W, int C, int Cpitch, int padded, cudaStream_t stream) {
  if (!x || !out) { ddg_set_last_error("pnhwc_to_nchw: bad args"); return DDG_ERR_ARG; }
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, N), block(32, 8);
  launch_pdl(pnhwc_to_nchw_kernel, grid, block, 0, stream, x, out, N, H, W, C, Cpitch, padded);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_gn_prepare(const double* stats_a, int Ca, const double* stats_b, int Cb, const float* gamma, const float* beta,
                              int gb_stride, int per_sample, float* scale, float* shift, int N, int HW, int G, float eps,
                              cudaStream_t stream) {
  const int C = Ca + (stats_b ? Cb : 0);
  if (!stats_a || !scale || !shift || G <= 0 || C % G != 0) { ddg_set_last_error("gn_prepare: bad args"); return DDG_ERR_ARG; }
  launch_pdl(gn_prepare_kernel, dim3(N), dim3(256), 2 * C * sizeof(double), stream, stats_a, Ca, stats_b, stats_b ? Cb : 0, gamma, beta, gb_stride,
             per_sample, scale, shift, HW, G, eps);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

namespace ddg {
// one thread per (sample, group); see ddg_gn_prepare_bwd in the header for the contract
__global__ void gn_prepare_bwd_kernel(const double* __restrict__ st, const float* __restrict__ gamma, int gb_stride, int per_sample,
                                      const float* __restrict__ dscale, const float* __restrict__ dshift, double* __restrict__ dst,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int C, int HW, int G, float eps) {
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
    const double ds = dscale[base + j], dh = dshift[base + j];
    const double t = ds - mean * dh;                 // scale = ga*rstd, shift = be - mean*ga*rstd
    dgamma[base + j] = (float)(rstd * t);
    dbeta[base + j] = (float)dh;
    d_rstd += ga * t;
    d_mean -= rstd * ga * dh;
  }
  const double d_var = clamped ? 0.0 : -0.5 * d_rstd * rstd * rstd * rstd;
  const double d_s1 = d_mean / cnt - 2.0 * mean * d_var / cnt;
  const double d_s2 = d_var / cnt;
  for (int j = 0; j < cpg; ++j) { dst[(base + j) * 2] = d_s1; dst[(base + j) * 2 + 1] = d_s2; }
}

__global__ void zero_border_kernel(float* __restrict__ buf, int N, int H, int W, int C4) {
  const int bp = 2 * (W + 2) + 2 * H;                // frame pixels per image
  const long total = (long)N * bp * C4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long r = i / C4;
    const int b = (int)(r % bp);
    const int n = (int)(r / bp);
    int hp, wp;
    if (b < W + 2) { hp = 0; wp = b; }
    else if (b < 2 * (W + 2)) { hp = H + 1; wp = b - (W + 2); }
    else { const int k = b - 2 * (W + 2); hp = 1 + (k >> 1); wp = (k & 1) ? W + 1 : 0; }
    reinterpret_cast<float4*>(buf + ((size_t)(n * (H + 2) + hp) * (W + 2) + wp) * (size_t)(C4 * 4))[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
}  // namespace ddg

extern "C" int ddg_gn_prepare_bwd(const double* stats, const float* gamma, int gb_stride, int per_sample, const float* dscale,
                                  const float* dshift, double* dstats, float* dgamma, float* dbeta, int N, int C, int HW, int G,
                                  float eps, cudaStream_t stream) {
  if (!stats || !dscale || !dshift || !dstats || !dgamma || !dbeta || G <= 0 || C % G != 0) { ddg_set_last_error("gn_prepare_bwd: bad args"); return DDG_ERR_ARG; }
  const int total = N * G;
  ddg::gn_prepare_bwd_kernel<<<(total + 127) / 128, 128, 0, stream>>>(stats, gamma, gb_stride, per_sample, dscale, dshift, dstats, dgamma,
                                                                    dbeta, N, C, HW, G, eps);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_zero_border(float* buf, int N, int H, int W, int C, cudaStream_t stream) {
  if (!buf || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 4 != 0) { ddg_set_last_error("zero_border: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * (2 * (W + 2) + 2 * H) * (C / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 8) blocks = 148L * 8;
  ddg::zero_border_kernel<<<(int)blocks, 256, 0, stream>>>(buf, N, H, W, C / 4);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_spatial_sum(const float* x, float* out, int N, int H, int W, int C, int act, cudaStream_t stream) {
  if (!x || !out) { ddg_set_last_error("spatial_sum: bad args"); return DDG_ERR_ARG; }
  dim3 grid((C + 127) / 128, N);
  spatial_sum_kernel<<<grid, 128, 0, stream>>>(x, out, H, W, C, act);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_minibatch_stddev(const float* x, float* out, int N, int H, int W, int C, int Cpad, int group, cudaStream_t stream) {
  if (!x || !out || group < 1 || group > 4 || N % group != 0) { ddg_set_last_error("minibatch_stddev: bad args"); return DDG_ERR_ARG; }
  minibatch_stddev_kernel<<<N / group, 256, 0, stream>>>(x, out, N, H, W, C, Cpad, group);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_channel_sum(const float* g, float* out, int N, int C, int inner, cudaStream_t stream) {
  if (!g || !out) { ddg_set_last_error("channel_sum: bad args"); return DDG_ERR_ARG; }
  channel_sum_kernel<<<C, 256, 0, stream>>>(g, out, N, C, inner);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

namespace ddg {
// one warp per row, T <= 32*32
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, float* __restrict__ p, long rows, int T, int lds,
                                                          int ldp) {
  const long row = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + row * lds;
  float* pr = p + row * ldp;
  float v[32];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = lane + 32 * i;
    v[i] = j < T ? sr[j] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = lane + 32 * i;
    v[i] = j < T ? expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = lane + 32 * i;
    if (j < ldp) pr[j] = j < T ? v[i] * inv : 0.f;
  }
}
}  // namespace ddg

extern "C" int ddg_softmax_rows(const float* s, float* p, long rows, int T, int lds, int ldp, cudaStream_t stream) {
  if (!s || !p || T <= 0 || T > 1024 || ldp > 1024 || lds < T || ldp < T) { ddg_set_last_error("softmax_rows: bad args (T <= 1024)"); return DDG_ERR_ARG; }
  ddg::softmax_rows_kernel<<<(int)((rows + 7) / 8), 256, 0, stream>>>(s, p, rows, T, lds, ldp);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

namespace ddg {
// backward of the row softmax: ds[r][j] = scale * p[r][j] * (dp[r][j] - sum_k p[r][k] dp[r][k]); one warp per row, T <= 1024
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ ds,
                                                              long rows, int T, int ld, float scale) {
  const long row = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* pr = p + row * ld;
  const float* gr = dp + row * ld;
  float* dr = ds + row * ld;
  float pv[32], gv[32];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = lane + 32 * i;
    pv[i] = j < T ? pr[j] : 0.f;
    gv[i] = j < T ? gr[j] : 0.f;
    dot += pv[i] * gv[i];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int j = lane + 32 * i;
    if (j < ld) dr[j] = j < T ? scale * pv[i] * (gv[i] - dot) : 0.f;
  }
}
}  // namespace ddg

extern "C" int ddg_softmax_rows_bwd(const float* p, const float* dp, float* ds, long rows, int T, int ld, float scale, cudaStream_t stream) {
  if (!p || !dp || !ds || T <= 0 || T > 1024 || ld > 1024 || ld < T) { ddg_set_last_error("softmax_rows_bwd: bad args (T <= 1024)"); return DDG_ERR_ARG; }
  ddg::softmax_rows_bwd_kernel<<<(int)((rows + 7) / 8), 256, 0, stream>>>(p, dp, ds, rows, T, ld, scale);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
