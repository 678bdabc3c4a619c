// fp16 / bf16 I/O variants of the two score_sde.op operators.  The reference dispatches both over
// AT_DISPATCH_FLOATING_TYPES_AND_HALF (upfirdn2d_kernel.cu:313, fused_bias_act_kernel.cu:79); here the fp32 entry points keep their
// specialised fast paths and the 16-bit types share two generic kernels: 16-bit loads / stores (8 elements = 16 bytes per
// access where alignment allows), fp32 taps, bias and accumulation.  Both are HBM-bound: 2 bytes per element each way.
#include <cuda_fp16.h>
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct LpParams {
  int in_h, in_w, out_h, out_w, kh, kw, up, down, pad0;
};

// upfirdn2d (upfirdn2d.py:184-225 semantics) on 16-bit planes.  A CTA owns a band of output rows of one plane: the input rows that
// band needs are brought into shared memory once as fp32 (16-byte loads of 8 elements), then every thread produces 8 consecutive
// outputs of one row (one 16-byte store) from shared memory.  UP / DOWN / KS are compile-time for the cases the models use (4x4
// taps, x2 up or down, or neither) so the zero-insertion index arithmetic folds away and the tap loops unroll; <0, 0, 0> is the
// general run-time form.
constexpr int kLpCap = 11 * 1024;   // floats of staged input per CTA (44 KB)

template <typename T, int UP, int DOWN, int KS>
__global__ void __launch_bounds__(256) upfirdn2d_lp_kernel(const T* __restrict__ x, const float* __restrict__ k, T* __restrict__ out,
                                                          LpParams p, int band_rows) {
  __shared__ float sk[256];        // flipped kernel
  __shared__ float sx[kLpCap];
  const int up = UP ? UP : p.up, down = DOWN ? DOWN : p.down, kh = KS ? KS : p.kh, kw = KS ? KS : p.kw;
  for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) {
    const int r = i / kw, c = i - r * kw;
    sk[i] = k[(kh - 1 - r) * kw + (kw - 1 - c)];
  }
  const long pl = blockIdx.x;
  const int oy0 = blockIdx.y * band_rows;
  const int oy1 = min(oy0 + band_rows, p.out_h);
  // input rows touched by output rows [oy0, oy1): a = oy*down - pad0 + i, iy = a / up
  int iy_lo = (oy0 * down - p.pad0) / up;
  if (oy0 * down - p.pad0 < 0) iy_lo = 0;
  int iy_hi = ((oy1 - 1) * down - p.pad0 + kh - 1) / up;
  iy_lo = max(iy_lo, 0);
  iy_hi = min(iy_hi, p.in_h - 1);
  const int nrows = max(iy_hi - iy_lo + 1, 0);
  const T* xin = x + pl * (long)p.in_h * p.in_w + (long)iy_lo * p.in_w;
  const int nelem = nrows * p.in_w;
  if ((p.in_w % 8 == 0) && ((((uintptr_t)xin) & 15) == 0)) {
    for (int i = threadIdx.x; i < nelem / 8; i += blockDim.x) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xin) + i);
      const T* v = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) sx[i * 8 + j] = to_f<T>(v[j]);
    }
  } else {
    for (int i = threadIdx.x; i < nelem; i += blockDim.x) sx[i] = to_f<T>(xin[i]);
  }
  __syncthreads();
  constexpr int VEC = 8;
  const int wv = (p.out_w + VEC - 1) / VEC;
  const int total = (oy1 - oy0) * wv;
  T* oplane = out + pl * (long)p.out_h * p.out_w;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int xv = idx % wv;
    const int oy = oy0 + idx / wv;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const int base_y = oy * down - p.pad0;
#pragma unroll(KS ? KS : 1)
    for (int i = 0; i < kh; ++i) {
      const int a = base_y + i;
      if (a < 0 || (a % up) != 0) continue;
      const int iy = a / up;
      if (iy >= p.in_h) continue;
      const float* row = sx + (iy - iy_lo) * p.in_w;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int base_x = (xv * VEC + v) * down - p.pad0;
#pragma unroll(KS ? KS : 1)
        for (int j = 0; j < kw; ++j) {
          const int b = base_x + j;
          if (b < 0 || (b % up) != 0) continue;
          const int ix = b / up;
          if (ix < p.in_w) acc[v] = fmaf(row[ix], sk[i * kw + j], acc[v]);
        }
      }
    }
    T* o = oplane + (long)oy * p.out_w + xv * VEC;
    if (xv * VEC + VEC <= p.out_w && ((((uintptr_t)o) & 15) == 0)) {
      T tmp[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) tmp[v] = from_f<T>(acc[v]);
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(tmp);
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (xv * VEC + v < p.out_w) o[v] = from_f<T>(acc[v]);
    }
  }
}

// ---- register-tiled fast paths for the two resampling shapes the models use (4x4 taps, in_w % 8 == 0, 16-byte aligned planes) ----
// Same scheme as the fp32 kernels of upfirdn2d.cu, at 8 elements (16 bytes) per load: no shared memory, one 16-byte load per input
// row per thread plus the two edge neighbours, fp32 taps / accumulation.
template <typename T>
__device__ __forceinline__ void lp_load_row8(const T* __restrict__ row, bool rok, bool has_l, bool has_r, float (&v)[10]) {
  uint4 raw = make_uint4(0u, 0u, 0u, 0u);
  if (rok) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(row));
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[1 + j] = to_f<T>(e[j]);
  v[0] = (rok && has_l) ? to_f<T>(__ldg(row - 1)) : 0.f;
  v[9] = (rok && has_r) ? to_f<T>(__ldg(row + 8)) : 0.f;
}

// down = 2, pad = (1, 1) (downsample_2d, up_or_down_sampling.py:240-257): thread = 4 output columns x 4 output rows
template <typename T>
__global__ void __launch_bounds__(256) lp_k4_down2_kernel(const T* __restrict__ x, const float* __restrict__ k, T* __restrict__ out, long planes,
                                                         int in_h, int in_w, int out_h, int out_w) {
  constexpr int RB = 4, NR = (RB - 1) * 2 + 4;
  float kf[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) kf[i][j] = __ldg(k + (3 - i) * 4 + (3 - j));
  const int hb = (out_h + RB - 1) / RB;
  const int wg = in_w >> 3;
  const long total = planes * hb * wg;
  const long in_plane = (long)in_h * in_w, out_plane = (long)out_h * out_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % wg);
    long r = idx / wg;
    const int oy0 = (int)(r % hb) * RB;
    const long pl = r / hb;
    const T* xin = x + pl * in_plane;
    const int by = oy0 * 2 - 1, x0 = 8 * g;
    float acc[RB][4];
#pragma unroll
    for (int q = 0; q < RB; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[q][c] = 0.f;
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
      const int iy = by + rr;
      const bool rok = (iy >= 0) && (iy < in_h);
      float v[10];   // columns x0-1 .. x0+8
      lp_load_row8<T>(xin + (long)(rok ? iy : 0) * in_w + x0, rok, x0 > 0, x0 + 8 < in_w, v);
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const int i = rr - q * 2;
        if (i >= 0 && i < 4) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            acc[q][c] = fmaf(v[2 * c], kf[i][0], fmaf(v[2 * c + 1], kf[i][1], fmaf(v[2 * c + 2], kf[i][2], fmaf(v[2 * c + 3], kf[i][3], acc[q][c]))));
        }
      }
    }
    T* o = out + pl * out_plane + (long)oy0 * out_w + 4 * g;
#pragma unroll
    for (int q = 0; q < RB; ++q)
      if (oy0 + q < out_h) {
        T pk[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) pk[c] = from_f<T>(acc[q][c]);
        *reinterpret_cast<uint2*>(o + (long)q * out_w) = *reinterpret_cast<const uint2*>(pk);
      }
  }
}

// up = 2, pad = (2, 1) (upsample_2d, up_or_down_sampling.py:213-237), polyphase: thread = 8 input columns of one input row ->
// 16 output columns x 2 output rows (two 32-byte stores)
template <typename T>
__global__ void __launch_bounds__(256) lp_k4_up2_kernel(const T* __restrict__ x, const float* __restrict__ k, T* __restrict__ out, long planes,
                                                       int in_h, int in_w) {
  float kf[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) kf[i][j] = __ldg(k + (3 - i) * 4 + (3 - j));
  const int wg = in_w >> 3;
  const long total = planes * in_h * wg;
  const int out_w = 2 * in_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % wg);
    long r = idx / wg;
    const int yy = (int)(r % in_h);
    const long pl = r / in_h;
    const int x0 = 8 * g;
    const T* xin = x + pl * (long)in_h * in_w;
    float v[3][10];   // rows yy-1 .. yy+1, columns x0-1 .. x0+8
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = yy + dy - 1;
      const bool rok = iy >= 0 && iy < in_h;
      lp_load_row8<T>(xin + (long)(rok ? iy : 0) * in_w + x0, rok, x0 > 0, x0 + 8 < in_w, v[dy]);
    }
    T* op = out + pl * (long)(4 * in_h * in_w) + (long)(2 * yy) * out_w + 2 * x0;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      T pk[16];
#pragma unroll
      for (int px = 0; px < 8; ++px)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float acc = 0.f;
#pragma unroll
          for (int ii = 0; ii < 2; ++ii)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int i = a + 2 * ii, j = b + 2 * jj;
              acc = fmaf(v[(a + i) / 2][px + (b + j) / 2], kf[i][j], acc);
            }
          pk[2 * px + b] = from_f<T>(acc);
        }
      uint4* dst = reinterpret_cast<uint4*>(op + (long)a * out_w);
      dst[0] = reinterpret_cast<const uint4*>(pk)[0];
      dst[1] = reinterpret_cast<const uint4*>(pk)[1];
    }
  }
}

template <typename T>
static int launch_upfirdn_lp(const void* x, const float* k, void* out, long planes, const LpParams& p, cudaStream_t stream) {
  // band of output rows per CTA such that the staged input rows fit: rows_in <= (band*down + kh) / up + 2
  const T* xi0 = (const T*)x;
  T* oo0 = (T*)out;
  const bool aligned = (p.in_w % 8 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)out & 15) == 0);
  if (aligned && p.kh == 4 && p.kw == 4 && p.up == 1 && p.down == 2 && p.pad0 == 1 && (p.in_h % 2 == 0) && p.out_h == p.in_h / 2 && p.out_w == p.in_w / 2) {
    const long total = planes * ((p.out_h + 3) / 4) * (p.in_w / 8);
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    lp_k4_down2_kernel<T><<<(int)blocks, 256, 0, stream>>>(xi0, k, oo0, planes, p.in_h, p.in_w, p.out_h, p.out_w);
    return DDG_OK;
  }
  if (aligned && p.kh == 4 && p.kw == 4 && p.up == 2 && p.down == 1 && p.pad0 == 2 && p.out_h == 2 * p.in_h && p.out_w == 2 * p.in_w) {
    const long total = planes * p.in_h * (p.in_w / 8);
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    lp_k4_up2_kernel<T><<<(int)blocks, 256, 0, stream>>>(xi0, k, oo0, planes, p.in_h, p.in_w);
    return DDG_OK;
  }
  const int rows_cap = kLpCap / p.in_w;
  int band = ((rows_cap - 2) * p.up - p.kh) / p.down;
  if (band < 1) return DDG_ERR_UNSUPPORTED;
  if (band > p.out_h) band = p.out_h;
  // enough CTAs to fill the machine when there are few planes
  while (band > 8 && planes * ((p.out_h + band - 1) / band) < 2 * 148) band = (band + 1) / 2;
  const dim3 grid((unsigned)planes, (unsigned)((p.out_h + band - 1) / band));
  const T* xi = (const T*)x;
  T* oo = (T*)out;
  if (p.kh == 4 && p.kw == 4 && p.up == 1 && p.down == 1) upfirdn2d_lp_kernel<T, 1, 1, 4><<<grid, 256, 0, stream>>>(xi, k, oo, p, band);
  else if (p.kh == 4 && p.kw == 4 && p.up == 1 && p.down == 2) upfirdn2d_lp_kernel<T, 1, 2, 4><<<grid, 256, 0, stream>>>(xi, k, oo, p, band);
  else if (p.kh == 4 && p.kw == 4 && p.up == 2 && p.down == 1) upfirdn2d_lp_kernel<T, 2, 1, 4><<<grid, 256, 0, stream>>>(xi, k, oo, p, band);
  else upfirdn2d_lp_kernel<T, 0, 0, 0><<<grid, 256, 0, stream>>>(xi, k, oo, p, band);
  return DDG_OK;
}

// fused_bias_act_kernel.cu:20-51 on 16-bit tensors (x, ref, y) with an fp32 bias; 8 elements per thread when step_b % 8 == 0
template <typename T>
__global__ void __launch_bounds__(256) fused_bias_act_lp_kernel(const T* __restrict__ x, const float* __restrict__ b, const T* __restrict__ ref,
                                                               T* __restrict__ y, long n, int step_b, int size_b, int act, int grad,
                                                               float alpha, float scale, int vec) {
  const long stride = (long)gridDim.x * blockDim.x;
  auto f = [&](float v, float r) {
    float o;
    if (act == 3) o = grad == 0 ? (v > 0 ? v : v * alpha) : (grad == 1 ? (r > 0 ? v : v * alpha) : 0.f);
    else o = grad == 2 ? 0.f : v;
    return o * scale;
  };
  if (vec) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n / 8; i += stride) {
      uint4 raw = *(reinterpret_cast<const uint4*>(x) + i);
      uint4 rr = make_uint4(0, 0, 0, 0);
      if (ref) rr = *(reinterpret_cast<const uint4*>(ref) + i);
      const T* xv = reinterpret_cast<const T*>(&raw);
      const T* rv = reinterpret_cast<const T*>(&rr);
      const float bb = b ? __ldg(b + ((i * 8) / step_b) % size_b) : 0.f;
      T o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = from_f<T>(f(to_f<T>(xv[j]) + bb, ref ? to_f<T>(rv[j]) : 0.f));
      *(reinterpret_cast<uint4*>(y) + i) = *reinterpret_cast<const uint4*>(o);
    }
  } else {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += stride) {
      float v = to_f<T>(x[i]);
      if (b) v += b[(i / step_b) % size_b];
      y[i] = from_f<T>(f(v, ref ? to_f<T>(ref[i]) : 0.f));
    }
  }
}

}  // namespace ddg

using namespace ddg;

extern "C" int ddg_upfirdn2d_lp(const void* x, const float* k, void* out, long planes, int in_h, int in_w, int kh, int kw, int up, int down,
                                int pad0, int pad1, int dtype, cudaStream_t stream) {
  if (!x || !k || !out || planes < 0 || in_h <= 0 || in_w <= 0 || kh <= 0 || kw <= 0 || up < 1 || down < 1 || (dtype != 1 && dtype != 2)) {
    ddg_set_last_error("upfirdn2d_lp: bad args (dtype 1 = fp16, 2 = bf16)");
    return DDG_ERR_ARG;
  }
  if (kh > 16 || kw > 16) { ddg_set_last_error("upfirdn2d_lp: kernel larger than 16 x 16 taps"); return DDG_ERR_UNSUPPORTED; }
  LpParams p;
  p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw; p.up = up; p.down = down; p.pad0 = pad0;
  p.out_h = (in_h * up + pad0 + pad1 - kh) / down + 1;
  p.out_w = (in_w * up + pad0 + pad1 - kw) / down + 1;
  if (p.out_h <= 0 || p.out_w <= 0) { ddg_set_last_error("upfirdn2d_lp: empty output"); return DDG_ERR_ARG; }
  if (planes == 0) return DDG_OK;
  if (planes > 2147483647L) { ddg_set_last_error("upfirdn2d_lp: too many planes"); return DDG_ERR_UNSUPPORTED; }
  const int rc = dtype == 1 ? launch_upfirdn_lp<__half>(x, k, out, planes, p, stream) : launch_upfirdn_lp<__nv_bfloat16>(x, k, out, planes, p, stream);
  if (rc != DDG_OK) { ddg_set_last_error("upfirdn2d_lp: image too wide for the staged kernel"); return rc; }
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_fused_bias_act_lp(const void* x, const float* b, const void* ref, void* y, long n, int step_b, int size_b, int act, int grad,
                                     float alpha, float scale, int dtype, cudaStream_t stream) {
  if (!x || !y || n < 0 || step_b < 1 || size_b < 1 || (dtype != 1 && dtype != 2)) { ddg_set_last_error("fused_bias_act_lp: bad args"); return DDG_ERR_ARG; }
  if (n == 0) return DDG_OK;
  const int vec = (step_b % 8 == 0 && n % 8 == 0 && ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)(ref ? ref : x))) & 15) == 0) ? 1 : 0;
  long blocks = ((vec ? n / 8 : n) + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  if (blocks < 1) blocks = 1;
  if (dtype == 1)
    fused_bias_act_lp_kernel<__half><<<(int)blocks, 256, 0, stream>>>((const __half*)x, b, (const __half*)ref, (__half*)y, n, step_b, size_b, act,
                                                                     grad, alpha, scale, vec);
  else
    fused_bias_act_lp_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)x, b, (const __nv_bfloat16*)ref,
                                                                            (__nv_bfloat16*)y, n, step_b, size_b, act, grad, alpha, scale, vec);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
