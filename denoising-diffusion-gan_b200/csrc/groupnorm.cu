// GroupNorm / AdaptiveGroupNorm (+ fused activation) on NCHW, forward and backward, and fused bias + leaky-ReLU.
// Replaces ATen native_group_norm as reached from layerspp.py:46-63,100 and ncsnpp_generator_adagn.py:264, and
// score_sde/op/fused_bias_act_kernel.cu:20-101.
//
// In NCHW the cpg channels of one (n, group) are one contiguous slab of cpg*HW floats.  Forward, main path
// (groupnorm_fwd_pipe_kernel): the slab is cut into <= 32 KB parts, one CTA per part, the CTAs of a slab forming a
// thread-block cluster (1..16).  Each part is brought in by 1-D bulk TMA behind an mbarrier, swept once in shared memory
// for shifted one-pass statistics, the parts' (mean, M2) are exchanged with st.async through distributed shared memory
// and merged (parallel-variance formula), and y = act(a*x + b) is written with streaming 128-bit stores from the same
// shared-memory copy: one HBM read + one HBM write, 8 B per element.  Slabs that do not fit (cluster of 16 x 104 KB) or
// are misaligned fall back to the one-CTA-per-slab kernels below.
#include <cstdlib>
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

// Per-channel coefficients of one (n, group) slab in shared memory: {a, b, a', b'} with y = act(a*x + b) and, for SiLU,
// the exponent argument -(a*x + b)*log2(e) = a'*x + b' prepared so that the inner loop is 2 FFMA + EX2 + ADD + RCP + MUL.
constexpr int kGnMaxCpg = 128;
__device__ __forceinline__ void gn_fill_coef(float4* coef, const float* gamma, const float* beta, int per_sample, int n, int C,
                                             int g, int cpg, float mean, float rstd) {
  for (int cl = threadIdx.x; cl < cpg; cl += blockDim.x) {
    const int c = g * cpg + cl;
    float ga = 1.f, be = 0.f;
    if (gamma) { ga = per_sample ? gamma[(size_t)n * C + c] : gamma[c]; be = per_sample ? beta[(size_t)n * C + c] : beta[c]; }
    const float a = ga * rstd, b = be - mean * a;
    coef[cl] = make_float4(a, b, -1.4426950408889634f * a, -1.4426950408889634f * b);
  }
}
template <int ACT>
__device__ __forceinline__ float gn_act(float x, const float4& k) {
  const float u = fmaf(x, k.x, k.y);
  if (ACT == ACT_SILU) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(x, k.z, k.w)));
    return __fdividef(u, 1.0f + e);
  }
  if (ACT == ACT_LEAKY) return u > 0.f ? u : 0.2f * u;
  if (ACT == ACT_TANH) return tanhf(u);
  return u;
}
// y4[k] = act(coef[channel(k)] applied to src[k]) for the len4 float4 of one slab; src is shared or global memory.
template <int ACT>
__device__ __forceinline__ void gn_apply(const float4* __restrict__ src, float4* __restrict__ y4, const float4* coef, int hw4, int len4, int koff) {
  const int sh = (hw4 & (hw4 - 1)) == 0 ? __ffs(hw4) - 1 : -1;
#pragma unroll 4
  for (int k = threadIdx.x; k < len4; k += blockDim.x) {
    const float4 kk = coef[sh >= 0 ? ((k + koff) >> sh) : ((k + koff) / hw4)];
    float4 v = src[k];
    v.x = gn_act<ACT>(v.x, kk); v.y = gn_act<ACT>(v.y, kk); v.z = gn_act<ACT>(v.z, kk); v.w = gn_act<ACT>(v.w, kk);
    stg_stream(y4 + k, v);
  }
}
__device__ __forceinline__ void gn_apply_dispatch(int act, const float4* src, float4* y4, const float4* coef, int hw4, int len4, int koff = 0) {
  switch (act) {
    case ACT_SILU: gn_apply<ACT_SILU>(src, y4, coef, hw4, len4, koff); break;
    case ACT_LEAKY: gn_apply<ACT_LEAKY>(src, y4, coef, hw4, len4, koff); break;
    case ACT_TANH: gn_apply<ACT_TANH>(src, y4, coef, hw4, len4, koff); break;
    default: gn_apply<ACT_NONE>(src, y4, coef, hw4, len4, koff); break;
  }
}

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
  if (vec && cpg <= kGnMaxCpg) {
    __shared__ float4 coef[kGnMaxCpg];
    gn_fill_coef(coef, gamma, beta, per_sample, n, C, g, cpg, mean, rstd);
    __syncthreads();
    gn_apply_dispatch(act, CACHED ? reinterpret_cast<const float4*>(cache) : reinterpret_cast<const float4*>(xg),
                      reinterpret_cast<float4*>(yg), coef, HW / 4, (int)(len / 4));
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

// Persistent, TMA-pipelined forward for groups that fit shared memory at least twice.  In NCHW one (sample, group) slab is
// cpg*HW contiguous floats, so a 1-D bulk copy brings it in; the slabs of the next groups stream in behind an mbarrier
// while this one is reduced (two smem passes: mean, then centred variance) and written, so HBM never idles between the
// load / reduce / store phases of one slab (the non-pipelined kernel reaches ~50% of copy bandwidth for that reason).
__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void gn_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nGN_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra GN_DONE;\nbra GN_WAIT;\nGN_DONE:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void gn_issue_slab(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  for (uint32_t o = 0; o < bytes; o += 32768u) {
    const uint32_t n = bytes - o < 32768u ? bytes - o : 32768u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + o),
                 "l"(reinterpret_cast<const char*>(src) + o), "r"(n), "r"(bar)
                 : "memory");
  }
}

// A slab larger than the pipeline's shared memory is split over a thread-block cluster: CTA r streams the r-th part,
// the parts' (mean, M2) are exchanged through distributed shared memory and merged with the parallel-variance formula.
constexpr int kGnMaxCluster = 16;
__global__ void __launch_bounds__(1024) groupnorm_fwd_pipe_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float* __restrict__ y,
                                                                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int C,
                                                                  int HW, int G, float eps, int per_sample, int act, int num_slabs,
                                                                  int nstage, int part_bytes, int cs, int num_clusters) {
  extern __shared__ __align__(128) unsigned char gsm[];
  __shared__ float2 red[32];
  __shared__ float2 exch[2][kGnMaxCluster];
  __shared__ __align__(8) unsigned long long bars[4];
  __shared__ __align__(8) unsigned long long xbar[2];
  __shared__ float4 coef[kGnMaxCpg];
  const int cpg = C / G;
  const int len = cpg * HW;                            // floats per slab
  const int plen = part_bytes / 4;                     // floats per part (len / cs)
  const int cid = blockIdx.x / cs, rank = blockIdx.x - cid * cs;
  const int mine = (num_slabs - cid + num_clusters - 1) / num_clusters;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gn_smem_u32(&bars[s])));
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gn_smem_u32(&xbar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // peers must see initialised exchange barriers before the first remote complete_tx
  if (cs > 1) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage && i < mine; ++i) {
      const int ng = cid + i * num_clusters;
      gn_issue_slab(gn_smem_u32(gsm + (size_t)i * part_bytes), x + (size_t)ng * len + (size_t)rank * plen, (uint32_t)part_bytes,
                    gn_smem_u32(&bars[i]));
    }
  }
  const int hw4 = HW / 4, plen4 = plen / 4;
  for (int i = 0; i < mine; ++i) {
    const int st = i % nstage;
    const int ng = cid + i * num_clusters;
    const int n = ng / G, g = ng - n * G;
    gn_mbar_wait(gn_smem_u32(&bars[st]), (uint32_t)(i / nstage) & 1u);
    const float4* c4 = reinterpret_cast<const float4*>(gsm + (size_t)st * part_bytes);
    // shifted one-pass statistics: with K = one sample of the part, d = x - K keeps sum(d^2) - sum(d)^2/len free of the
    // mean^2 cancellation of the raw moments, at a single sweep over shared memory
    const float K = reinterpret_cast<const float*>(c4)[0];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int k = threadIdx.x; k < plen4; k += blockDim.x) {
      const float4 v = c4[k];
      const float a = v.x - K, b = v.y - K, c = v.z - K, d = v.w - K;
      s1 += (a + b) + (c + d);
      s2 += (a * a + b * b) + (c * c + d * d);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) red[warp] = make_float2(s1, s2);
    __syncthreads();
    const float2 r = lane < nwarp ? red[lane] : make_float2(0.f, 0.f);   // every warp reduces the warp partials itself
    const float t1 = warp_sum(r.x), t2 = warp_sum(r.y);
    const float md = t1 / (float)plen;
    float mean = K + md;
    float m2 = fmaxf(t2 - t1 * md, 0.f);
    if (cs > 1) {
      // exchange through st.async + the receiver's mbarrier: no release fence, so the y stores of the previous slab still in
      // flight are not drained on the critical path (a barrier.cluster release here costs a full store round trip per slab)
      const uint32_t xb = gn_smem_u32(&xbar[i & 1]);
      if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xb), "r"((uint32_t)cs * 8u) : "memory");
      if (threadIdx.x < cs) {
        uint32_t rdata, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdata) : "r"(gn_smem_u32(&exch[i & 1][rank])), "r"((uint32_t)threadIdx.x));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(xb), "r"((uint32_t)threadIdx.x));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(rdata), "f"(mean),
                     "f"(m2), "r"(rbar)
                     : "memory");
      }
      gn_mbar_wait(xb, (uint32_t)(i >> 1) & 1u);
      float msum = 0.f;
      for (int q = 0; q < cs; ++q) msum += exch[i & 1][q].x;
      const float mall = msum / (float)cs;
      float m2all = 0.f;
      for (int q = 0; q < cs; ++q) { const float2 e = exch[i & 1][q]; const float dm = e.x - mall; m2all += e.y + (float)plen * dm * dm; }
      mean = mall; m2 = m2all;
    }
    const float rstd = rsqrtf(m2 / (float)len + eps);
    if (threadIdx.x == 0 && rank == 0) {
      if (mean_out) mean_out[ng] = mean;
      if (rstd_out) rstd_out[ng] = rstd;
    }
    gn_fill_coef(coef, gamma, beta, per_sample, n, C, g, cpg, mean, rstd);
    __syncthreads();
    gn_apply_dispatch(act, c4, reinterpret_cast<float4*>(y + (size_t)ng * len + (size_t)rank * plen), coef, hw4, plen4, rank * plen4);
    __syncthreads();                                   // every thread is done reading this stage, red[] and coef[]
    if (threadIdx.x == 0 && i + nstage < mine) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const int ng2 = cid + (i + nstage) * num_clusters;
      gn_issue_slab(gn_smem_u32(gsm + (size_t)st * part_bytes), x + (size_t)ng2 * len + (size_t)rank * plen, (uint32_t)part_bytes,
                    gn_smem_u32(&bars[st]));
    }
  }
  // no CTA may leave while a peer can still store into its exchange slots
  if (cs > 1) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
  // One CTA per (n, group) slab.  Pass 1 sweeps the slab channel by channel with 128-bit loads; every warp leaves its partial
  // {sum du*xhat, sum du} per channel in shared memory and ONE barrier later the first threads finish the per-channel sums (the
  // first version paid six barriers per channel).  Pass 2 re-reads the slab (it is L2-resident: <= a few hundred KB) for dx.
  __shared__ float part[kGnMaxCpg][2][kGnThreads / 32];
  __shared__ float chan[kGnMaxCpg][4];      // per channel: a = sum du*xhat, b = sum du, gamma, beta
  __shared__ float msum[2];
  const int ng = blockIdx.x;
  const int n = ng / G, g = ng - n * G;
  const int cpg = C / G;
  const long len = (long)cpg * HW;
  const size_t base = ((size_t)n * C + (size_t)g * cpg) * HW;
  const float mean = mean_in[ng], rstd = rstd_in[ng];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (HW % 4 == 0) && ((base % 4) == 0);
  for (int cl = threadIdx.x; cl < cpg; cl += blockDim.x) {
    const int c = g * cpg + cl;
    float ga = 1.f, be = 0.f;
    if (gamma) { ga = per_sample ? gamma[(size_t)n * C + c] : gamma[c]; be = per_sample ? beta[(size_t)n * C + c] : beta[c]; }
    chan[cl][2] = ga; chan[cl][3] = be;
  }
  __syncthreads();
  for (int cl = 0; cl < cpg; ++cl) {
    const float ga = chan[cl][2], be = chan[cl][3];
    const float* xc = x + base + (size_t)cl * HW;
    const float* gc = dy + base + (size_t)cl * HW;
    float a = 0.f, b = 0.f;
    if (vec) {
      for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(xc) + i);
        const float4 gv = ldg_stream(reinterpret_cast<const float4*>(gc) + i);
        float xh, du;
        xh = (xv.x - mean) * rstd; du = gv.x * act_grad(fmaf(ga, xh, be), act); a += du * xh; b += du;
        xh = (xv.y - mean) * rstd; du = gv.y * act_grad(fmaf(ga, xh, be), act); a += du * xh; b += du;
        xh = (xv.z - mean) * rstd; du = gv.z * act_grad(fmaf(ga, xh, be), act); a += du * xh; b += du;
        xh = (xv.w - mean) * rstd; du = gv.w * act_grad(fmaf(ga, xh, be), act); a += du * xh; b += du;
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float xh = (xc[i] - mean) * rstd;
        const float du = gc[i] * act_grad(fmaf(ga, xh, be), act);
        a += du * xh;
        b += du;
      }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { part[cl][0][warp] = a; part[cl][1][warp] = b; }
  }
  __syncthreads();
  for (int cl = threadIdx.x; cl < cpg; cl += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < kGnThreads / 32; ++w) { a += part[cl][0][w]; b += part[cl][1][w]; }
    chan[cl][0] = a; chan[cl][1] = b;
    const int c = g * cpg + cl;
    if (dgamma) dgamma[(size_t)n * C + c] = a;
    if (dbeta) dbeta[(size_t)n * C + c] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s1 = 0.f, s2 = 0.f;
    for (int cl = 0; cl < cpg; ++cl) { s1 += chan[cl][1] * chan[cl][2]; s2 += chan[cl][0] * chan[cl][2]; }
    msum[0] = s1 / (float)len; msum[1] = s2 / (float)len;
  }
  __syncthreads();
  const float m1 = msum[0], m2 = msum[1];
  if (vec) {
    const int hw4 = HW / 4;
    for (long i = threadIdx.x; i < len / 4; i += blockDim.x) {
      const int cl = (int)(i / hw4);
      const float ga = chan[cl][2], be = chan[cl][3];
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base) + i);
      const float4 gv = ldg_stream(reinterpret_cast<const float4*>(dy + base) + i);
      float4 o;
      float xh, du;
      xh = (xv.x - mean) * rstd; du = gv.x * act_grad(fmaf(ga, xh, be), act); o.x = rstd * (du * ga - m1 - xh * m2);
      xh = (xv.y - mean) * rstd; du = gv.y * act_grad(fmaf(ga, xh, be), act); o.y = rstd * (du * ga - m1 - xh * m2);
      xh = (xv.z - mean) * rstd; du = gv.z * act_grad(fmaf(ga, xh, be), act); o.z = rstd * (du * ga - m1 - xh * m2);
      xh = (xv.w - mean) * rstd; du = gv.w * act_grad(fmaf(ga, xh, be), act); o.w = rstd * (du * ga - m1 - xh * m2);
      stg_stream(reinterpret_cast<float4*>(dx + base) + i, o);
    }
  } else {
    for (long i = threadIdx.x; i < len; i += blockDim.x) {
      const int cl = (int)(i / HW);
      const float ga = chan[cl][2], be = chan[cl][3];
      const float xh = (x[base + i] - mean) * rstd;
      const float du = dy[base + i] * act_grad(fmaf(ga, xh, be), act);
      dx[base + i] = rstd * (du * ga - m1 - xh * m2);
    }
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
  const bool aligned = (HW % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
  if (aligned && bytes >= 2048 && C / G <= kGnMaxCpg) {
    // cluster size: smallest power of two whose part fits the pipeline twice
    int cs = 1;
    // part size: 32 KB (many small CTAs per SM) up to 256 KB slabs; 128 KB single-buffered parts for larger slabs, where the
    // per-slab exchange across a 16-CTA cluster would cost as much as moving a 64 KB part (measured 0.52 -> 0.62 of copy peak)
    static long target_env = -1;
    if (target_env < 0) { const char* e = getenv("DDG_GN_PART_KB"); target_env = e ? atol(e) * 1024 : 0; }
    const size_t target = target_env > 0 ? (size_t)target_env : (bytes >= 512 * 1024 ? 128 * 1024 : 32 * 1024);
    while (cs < kGnMaxCluster && (bytes / cs) > target) cs *= 2;
    const size_t part = bytes / cs;
    if (part <= 200 * 1024 && bytes % ((size_t)cs * 16) == 0) {
      static bool attr_p = false;
      if (!attr_p) {
        cudaFuncSetAttribute(groupnorm_fwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
        cudaFuncSetAttribute(groupnorm_fwd_pipe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        attr_p = true;
      }
      // measured (profiles/r1_microbench_sweep.txt): parts of <= 64 KB run best single-staged with many CTAs per SM
      // (their load / reduce / store phases interleave across CTAs); larger parts keep a second stage in flight
      static int ns_env = -1;
      if (ns_env < 0) { const char* e = getenv("DDG_GN_NSTAGE"); ns_env = e ? atoi(e) : 0; }
      int nstage = (part <= 64 * 1024 || part > 104 * 1024) ? 1 : 2;
      if (ns_env >= 1 && ns_env <= 4 && (size_t)ns_env * part <= 208 * 1024) nstage = ns_env;
      int threads = 128;
      while (threads < 1024 && (size_t)threads * 64 < part) threads *= 2;   // >= 4 float4 per thread per sweep
      // small parts: several CTAs per SM instead of deeper pipelines
      int per_sm = (int)((208 * 1024) / (nstage * part));
      if (per_sm > 2048 / threads) per_sm = 2048 / threads;
      if (per_sm > 8) per_sm = 8;
      if (per_sm < 1) per_sm = 1;
      long num_clusters = 148L * per_sm / cs;
      if (num_clusters > (long)N * G) num_clusters = (long)N * G;
      if (num_clusters < 1) num_clusters = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(num_clusters * cs));
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = nstage * part;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cs > 1) {
        // a persistent grid must be co-resident: GPC granularity leaves fewer cluster slots than SMs / cs
        static int cached_key = -1, cached_max = 0;
        const int key = cs * 4096 + threads + (int)(cfg.dynamicSmemBytes >> 10) * 65536;
        if (key != cached_key) {
          int mc = 0;
          if (cudaOccupancyMaxActiveClusters(&mc, groupnorm_fwd_pipe_kernel, &cfg) != cudaSuccess || mc < 1) mc = 1;
          cached_key = key; cached_max = mc;
        }
        if (num_clusters > cached_max) { num_clusters = cached_max; cfg.gridDim = dim3((unsigned)(num_clusters * cs)); }
      }
      cudaError_t e = cudaLaunchKernelEx(&cfg, groupnorm_fwd_pipe_kernel, x, gamma, beta, y, mean, rstd, C, HW, G, eps, per_sample, act,
                                         N * G, nstage, (int)part, cs, (int)num_clusters);
      if (e != cudaSuccess) { ddg_set_last_error(cudaGetErrorString(e)); return DDG_ERR_LAUNCH; }
      return DDG_OK;
    }
  }
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
