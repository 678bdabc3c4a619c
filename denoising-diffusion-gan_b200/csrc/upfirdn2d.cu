// upfirdn2d for sm_100a: pad -> zero-insert up -> FIR (true convolution) -> decimate.
// Replaces score_sde/op/upfirdn2d_kernel.cu:51-371 behind the same argument list (upfirdn2d.cpp:20-31).
//
// The reference launches one 256-thread CTA per 16x64 / 8x32 output tile per (N*C) plane, which is 6-50 % filled on
// the <= 32 px maps of the CIFAR config.  Here a plane is a contiguous run of memory, the
// grid is a flat grid-stride loop over float4 groups of outputs (HBM-bound op: bytes = 4*(in + out) per plane) and
// only the polyphase taps that hit a non-zero sample are visited (4 of 16 for up=2).
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

constexpr int kMaxTaps = 256;

struct UpfirdnParams {
  int in_h, in_w, out_h, out_w, kh, kw;
  int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
};

// generic: VEC consecutive outputs along x per thread
template <int VEC>
__global__ void __launch_bounds__(256) upfirdn2d_kernel(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ out,
                                                        long planes, UpfirdnParams p) {
  __shared__ float sk[kMaxTaps];  // flipped kernel: sk[i][j] = k[kh-1-i][kw-1-j]
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    const int r = i / p.kw, c = i - r * p.kw;
    sk[i] = k[(p.kh - 1 - r) * p.kw + (p.kw - 1 - c)];
  }
  __syncthreads();
  const int wv = (p.out_w + VEC - 1) / VEC;
  const long total = planes * p.out_h * wv;
  const long in_plane = (long)p.in_h * p.in_w, out_plane = (long)p.out_h * p.out_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int xv = (int)(idx % wv);
    long r = idx / wv;
    const int oy = (int)(r % p.out_h);
    const long pl = r / p.out_h;
    const float* xin = x + pl * in_plane;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const int base_y = oy * p.down_y - p.pad_y0;
    int i0 = (-base_y) % p.up_y;
    if (i0 < 0) i0 += p.up_y;
    for (int i = i0; i < p.kh; i += p.up_y) {
      const int a = base_y + i;
      if (a < 0) continue;
      const int iy = a / p.up_y;
      if (iy >= p.in_h) break;
      const float* row = xin + (long)iy * p.in_w;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int ox = xv * VEC + v;
        const int base_x = ox * p.down_x - p.pad_x0;
        int j0 = (-base_x) % p.up_x;
        if (j0 < 0) j0 += p.up_x;
        for (int j = j0; j < p.kw; j += p.up_x) {
          const int b = base_x + j;
          if (b < 0) continue;
          const int ix = b / p.up_x;
          if (ix >= p.in_w) break;
          acc[v] = fmaf(__ldg(row + ix), sk[i * p.kw + j], acc[v]);
        }
      }
    }
    float* o = out + pl * out_plane + (long)oy * p.out_w + xv * VEC;
    if (VEC == 4 && xv * 4 + 3 < p.out_w && ((((uintptr_t)o) & 15) == 0)) {
      stg_stream(reinterpret_cast<float4*>(o), make_float4(acc[0], acc[1], acc[2], acc[3]));
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (xv * VEC + v < p.out_w) o[v] = acc[v];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fast paths for 4x4-tap kernels (the only kernel size the models use: [1,3,3,1] (x) [1,3,3,1]):
//   <UP=1, DOWN=1|2>: each thread produces 4 adjacent outputs of one row from a 4 x (3*DOWN+4) register patch
//                     (10 loads per output row instead of 16 per output), taps in registers, predicated loads.
//   <UP=2, DOWN=1>  : polyphase -- every output uses 2x2 of the 16 taps; 4 adjacent outputs share a 2 x 4 input patch.
// Flat grid-stride indexing over (plane, row, column group) keeps every thread busy on the 4..64-pixel planes of the CIFAR
// configuration (the reference's 16x64 / 8x32 tiles are 6-50 % filled there).
// ---------------------------------------------------------------------------------------------------------
template <int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_k4_down_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                               float* __restrict__ out, long planes, int in_h, int in_w, int out_h,
                                                               int out_w, int pad_x0, int pad_y0) {
  // adjacent lanes = adjacent output columns (coalesced loads/stores); each thread walks RB output rows so that the
  // 4-column input patch of a row is loaded once and reused by every output row it contributes to.
  constexpr int RB = 4;
  constexpr int NR = (RB - 1) * DOWN + 4;
  float kf[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) kf[i][j] = __ldg(k + (3 - i) * 4 + (3 - j));   // flipped: true convolution
  const int hb = (out_h + RB - 1) / RB;
  const long total = planes * hb * out_w;
  const long in_plane = (long)in_h * in_w, out_plane = (long)out_h * out_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % out_w);
    long r = idx / out_w;
    const int oy0 = (int)(r % hb) * RB;
    const long pl = r / hb;
    const float* xin = x + pl * in_plane;
    const int by = oy0 * DOWN - pad_y0, bx = ox * DOWN - pad_x0;
    float acc[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) acc[q] = 0.f;
    const bool c0 = bx >= 0 && bx < in_w, c1 = bx + 1 >= 0 && bx + 1 < in_w, c2 = bx + 2 >= 0 && bx + 2 < in_w,
               c3 = bx + 3 >= 0 && bx + 3 < in_w;
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
      const int iy = by + rr;
      const bool rok = (iy >= 0) && (iy < in_h);
      const float* row = xin + (long)(rok ? iy : 0) * in_w + bx;
      const float v0 = (rok && c0) ? __ldg(row) : 0.f;
      const float v1 = (rok && c1) ? __ldg(row + 1) : 0.f;
      const float v2 = (rok && c2) ? __ldg(row + 2) : 0.f;
      const float v3 = (rok && c3) ? __ldg(row + 3) : 0.f;
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const int i = rr - q * DOWN;               // tap row of output row q fed by input row rr (compile time)
        if (i >= 0 && i < 4) acc[q] = fmaf(v0, kf[i][0], fmaf(v1, kf[i][1], fmaf(v2, kf[i][2], fmaf(v3, kf[i][3], acc[q]))));
      }
    }
    float* o = out + pl * out_plane + (long)oy0 * out_w + ox;
#pragma unroll
    for (int q = 0; q < RB; ++q)
      if (oy0 + q < out_h) o[(long)q * out_w] = acc[q];
  }
}

// up = 2, pad0 = 2 (upsample_2d's pad (2,1)): out[2y+a][2x+b] = sum_{i in {a,a+2}} sum_{j in {b,b+2}}
//   in[y + (a+i)/2 - 1][x + (b+j)/2 - 1] * kf[i][j].  One thread per input pixel: 3x3 patch in, 2x2 block out, all 16 taps used
// exactly once with compile-time indices; lanes = adjacent x (coalesced 128 B loads, 256 B stores).
__global__ void __launch_bounds__(256) upfirdn2d_k4_up2_poly_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                                   float* __restrict__ out, long planes, int in_h, int in_w) {
  float kf[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) kf[i][j] = __ldg(k + (3 - i) * 4 + (3 - j));
  const long total = planes * in_h * in_w;
  const int out_w = 2 * in_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int xx = (int)(idx % in_w);
    long r = idx / in_w;
    const int yy = (int)(r % in_h);
    const long pl = r / in_h;
    const float* xin = x + pl * (long)in_h * in_w;
    float v[3][3];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = yy + dy - 1;
      const bool rok = iy >= 0 && iy < in_h;
      const float* row = xin + (long)(rok ? iy : 0) * in_w + xx;
      v[dy][0] = (rok && xx > 0) ? __ldg(row - 1) : 0.f;
      v[dy][1] = rok ? __ldg(row) : 0.f;
      v[dy][2] = (rok && xx + 1 < in_w) ? __ldg(row + 1) : 0.f;
    }
    float o[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float acc = 0.f;
#pragma unroll
        for (int ii = 0; ii < 2; ++ii)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int i = a + 2 * ii, j = b + 2 * jj;
            acc = fmaf(v[(a + i) / 2][(b + j) / 2], kf[i][j], acc);
          }
        o[a][b] = acc;
      }
    float* op = out + pl * (long)(4 * in_h * in_w) + (long)(2 * yy) * out_w + 2 * xx;
    *reinterpret_cast<float2*>(op) = make_float2(o[0][0], o[0][1]);
    *reinterpret_cast<float2*>(op + out_w) = make_float2(o[1][0], o[1][1]);
  }
}

__global__ void __launch_bounds__(256) upfirdn2d_k4_up2_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                              float* __restrict__ out, long planes, int in_h, int in_w, int out_h,
                                                              int out_w, int pad_x0, int pad_y0) {
  __shared__ float sk[16];
  if (threadIdx.x < 16) sk[threadIdx.x] = k[(3 - threadIdx.x / 4) * 4 + (3 - threadIdx.x % 4)];
  __syncthreads();
  const int wv = (out_w + 3) >> 2;
  const long total = planes * out_h * wv;
  const long in_plane = (long)in_h * in_w, out_plane = (long)out_h * out_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int xv = (int)(idx % wv);
    long r = idx / wv;
    const int oy = (int)(r % out_h);
    const long pl = r / out_h;
    const float* xin = x + pl * in_plane;
    const int by = oy - pad_y0, bx = xv * 4 - pad_x0;
    // rows: a = by + i even -> i in {i0, i0 + 2}, input row (by + i) >> 1
    const int i0 = by & 1;
    const int ixb = (bx + (bx & 1)) >> 1;             // first input column touched by the 4 outputs
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int i = i0 + 2 * ii;
      const int iy = (by + i) >> 1;
      const bool rok = (by + i >= 0) && (iy < in_h);
      const float* row = xin + (long)(rok ? iy : 0) * in_w;
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ix = ixb + c;
        v[c] = (rok && ix >= 0 && ix < in_w) ? __ldg(row + ix) : 0.f;
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int b = bx + o;
        const int j0 = b & 1;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = j0 + 2 * jj;
          const int c = ((b + j) >> 1) - ixb;          // 0..3
          const float xv_ = c == 0 ? v[0] : (c == 1 ? v[1] : (c == 2 ? v[2] : v[3]));
          acc[o] = fmaf(xv_, sk[i * 4 + j], acc[o]);
        }
      }
    }
    float* o = out + pl * out_plane + (long)oy * out_w + xv * 4;
    if (xv * 4 + 3 < out_w && ((((uintptr_t)o) & 15) == 0)) {
      stg_stream(reinterpret_cast<float4*>(o), make_float4(acc[0], acc[1], acc[2], acc[3]));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (xv * 4 + q < out_w) o[q] = acc[q];
    }
  }
}

// Vectorised variants (in_w % 4 == 0): one 128-bit load per input row per thread keeps >= 16 unique DRAM bytes in flight per
// thread (the scalar versions above are memory-level-parallelism bound at ~45 % of the HBM roofline).
__global__ void __launch_bounds__(256) upfirdn2d_k4_up2_poly4_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                                    float* __restrict__ out, long planes, int in_h, int in_w) {
  float kf[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) kf[i][j] = __ldg(k + (3 - i) * 4 + (3 - j));
  const int wq = in_w >> 2;
  const long total = planes * in_h * wq;
  const int out_w = 2 * in_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int xq = (int)(idx % wq);
    long r = idx / wq;
    const int yy = (int)(r % in_h);
    const long pl = r / in_h;
    const int x0 = xq * 4;
    const float* xin = x + pl * (long)in_h * in_w;
    float v[3][6];   // columns x0-1 .. x0+4
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = yy + dy - 1;
      const bool rok = iy >= 0 && iy < in_h;
      const float* row = xin + (long)(rok ? iy : 0) * in_w + x0;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rok) c = ldg_stream(reinterpret_cast<const float4*>(row));
      v[dy][0] = (rok && x0 > 0) ? __ldg(row - 1) : 0.f;
      v[dy][1] = c.x; v[dy][2] = c.y; v[dy][3] = c.z; v[dy][4] = c.w;
      v[dy][5] = (rok && x0 + 4 < in_w) ? __ldg(row + 4) : 0.f;
    }
    float o[2][8];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int a = 0; a < 2; ++a)
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
          o[a][2 * px + b] = acc;
        }
    float* op = out + pl * (long)(4 * in_h * in_w) + (long)(2 * yy) * out_w + 2 * x0;
    stg_stream(reinterpret_cast<float4*>(op), make_float4(o[0][0], o[0][1], o[0][2], o[0][3]));
    stg_stream(reinterpret_cast<float4*>(op) + 1, make_float4(o[0][4], o[0][5], o[0][6], o[0][7]));
    stg_stream(reinterpret_cast<float4*>(op + out_w), make_float4(o[1][0], o[1][1], o[1][2], o[1][3]));
    stg_stream(reinterpret_cast<float4*>(op + out_w) + 1, make_float4(o[1][4], o[1][5], o[1][6], o[1][7]));
  }
}

// down = 2, pad0 = 1 (downsample_2d): thread = 2 adjacent output columns x 4 output rows; per input row one float4
// (columns 4g .. 4g+3) plus the two neighbours 4g-1 and 4g+4.
__global__ void __launch_bounds__(256) upfirdn2d_k4_down2_vec_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                                    float* __restrict__ out, long planes, int in_h, int in_w, int out_h,
                                                                    int out_w) {
  constexpr int RB = 4, NR = (RB - 1) * 2 + 4;
  float kf[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) kf[i][j] = __ldg(k + (3 - i) * 4 + (3 - j));
  const int hb = (out_h + RB - 1) / RB;
  const int wh = out_w >> 1;
  const long total = planes * hb * wh;
  const long in_plane = (long)in_h * in_w, out_plane = (long)out_h * out_w;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % wh);
    long r = idx / wh;
    const int oy0 = (int)(r % hb) * RB;
    const long pl = r / hb;
    const float* xin = x + pl * in_plane;
    const int by = oy0 * 2 - 1, x0 = 4 * g;
    float acc[RB][2];
#pragma unroll
    for (int q = 0; q < RB; ++q) acc[q][0] = acc[q][1] = 0.f;
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
      const int iy = by + rr;
      const bool rok = (iy >= 0) && (iy < in_h);
      const float* row = xin + (long)(rok ? iy : 0) * in_w + x0;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rok) c = ldg_stream(reinterpret_cast<const float4*>(row));
      const float vl = (rok && x0 > 0) ? __ldg(row - 1) : 0.f;
      const float vr = (rok && x0 + 4 < in_w) ? __ldg(row + 4) : 0.f;
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const int i = rr - q * 2;
        if (i >= 0 && i < 4) {
          acc[q][0] = fmaf(vl, kf[i][0], fmaf(c.x, kf[i][1], fmaf(c.y, kf[i][2], fmaf(c.z, kf[i][3], acc[q][0]))));
          acc[q][1] = fmaf(c.y, kf[i][0], fmaf(c.z, kf[i][1], fmaf(c.w, kf[i][2], fmaf(vr, kf[i][3], acc[q][1]))));
        }
      }
    }
    float* o = out + pl * out_plane + (long)oy0 * out_w + 2 * g;
#pragma unroll
    for (int q = 0; q < RB; ++q)
      if (oy0 + q < out_h) *reinterpret_cast<float2*>(o + (long)q * out_w) = make_float2(acc[q][0], acc[q][1]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// FIR on the internal PNHWC layout, [1,3,3,1] (x) [1,3,3,1] / 64 (up: x4 gain), AdaGN + activation fused on load.
// One thread per (output pixel, 4 channels).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load_tr(const float* __restrict__ x, size_t off, const float4& sc, const float4& sh, bool affine,
                                          int act) {
  float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
  if (affine) { v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w); }
  if (act != ACT_NONE) { v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act); }
  return v;
}
__device__ __forceinline__ void fma4(float4& a, const float4& v, float w) {
  a.x = fmaf(v.x, w, a.x); a.y = fmaf(v.y, w, a.y); a.z = fmaf(v.z, w, a.z); a.w = fmaf(v.w, w, a.w);
}

// mode 1: up2 (H -> 2H), mode 2: down2 (H -> H/2), mode 3: pad(2,2) FIR H -> H+1 stored space-to-depth
__global__ void __launch_bounds__(256) fir_pnhwc_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int act, float* __restrict__ out, int N, int H,
                                                        int W, int C, int mode, int out_pitch, float gain) {
  pdl_wait();
  pdl_trigger();
  const float t4[4] = {1.f, 3.f, 3.f, 1.f};
  const int C4 = C / 4;
  int OH, OW;
  // (H, W) is always the size of the *non-s2d* image: the input for modes 1-3, the output for mode 4
  if (mode == 1) { OH = 2 * H; OW = 2 * W; } else if (mode == 2) { OH = H / 2; OW = W / 2; } else if (mode == 3) { OH = H + 1; OW = W + 1; }
  else { OH = H; OW = W; }
  const long total = (long)N * OH * OW * C4;
  const bool affine = scale != nullptr;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    long r = idx / C4;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
    if (affine) {
      sc = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * C) + c4);
      sh = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * C) + c4);
    }
    float4 acc = make_float4(0, 0, 0, 0);
    // In every mode out[oy][ox] = sum_{i,j} in[(oy*down + i - pad)/up] * kflip[i][j] over contributing taps; the
    // separable [1,3,3,1] kernel is symmetric so kflip == k.
    if (mode == 1) {
      // up=2, pad0=2: a = oy + i - 2 must be even and in range; weights k*4/64
      for (int i = (oy & 1); i < 4; i += 2) {
        const int iy = (oy + i - 2) >> 1;
        if (oy + i - 2 < 0 || iy >= H) continue;
        for (int j = (ox & 1); j < 4; j += 2) {
          const int ix = (ox + j - 2) >> 1;
          if (ox + j - 2 < 0 || ix >= W) continue;
          const size_t off = ((size_t)(n * (H + 2) + iy + 1) * (W + 2) + (ix + 1)) * C + c4 * 4;
          fma4(acc, load_tr(x, off, sc, sh, affine, act), t4[i] * t4[j] * (4.f / 64.f) * gain);
        }
      }
    } else if (mode == 4) {
      // adjoint of mode 3: dX[y][x] = sum_{i,j} k[i][j] * gfir[y - i + 2][x - j + 2], gfir stored space-to-depth with pitch
      // out_pitch (= the s2d tensor's channel pitch, 4*C) in cells of (H/2 + 3) x (W/2 + 3)
      const int Hc = H / 2 + 3, Wc = W / 2 + 3;
      for (int i = 0; i < 4; ++i) {
        const int fy = oy - i + 2;
        if (fy < 0 || fy > H) continue;
        for (int j = 0; j < 4; ++j) {
          const int fx = ox - j + 2;
          if (fx < 0 || fx > W) continue;
          const size_t off = ((size_t)(n * Hc + (fy >> 1) + 1) * Wc + ((fx >> 1) + 1)) * out_pitch + (size_t)((fy & 1) * 2 + (fx & 1)) * C + c4 * 4;
          fma4(acc, __ldg(reinterpret_cast<const float4*>(x + off)), t4[i] * t4[j] * (1.f / 64.f) * gain);
        }
      }
    } else {
      const int down = (mode == 2) ? 2 : 1;
      const int pad = (mode == 2) ? 1 : 2;
      for (int i = 0; i < 4; ++i) {
        const int iy = oy * down + i - pad;
        if (iy < 0 || iy >= H) continue;
        for (int j = 0; j < 4; ++j) {
          const int ix = ox * down + j - pad;
          if (ix < 0 || ix >= W) continue;
          const size_t off = ((size_t)(n * (H + 2) + iy + 1) * (W + 2) + (ix + 1)) * C + c4 * 4;
          fma4(acc, load_tr(x, off, sc, sh, affine, act), t4[i] * t4[j] * (1.f / 64.f) * gain);
        }
      }
    }
    size_t ooff;
    if (mode == 3) {
      // space-to-depth cell (oy/2, ox/2), sub-position (oy&1, ox&1); buffer [N][Ho+3][Wo+3][out_pitch], Ho = H/2
      const int Hc = H / 2 + 3, Wc = W / 2 + 3;
      ooff = ((size_t)(n * Hc + (oy >> 1) + 1) * Wc + ((ox >> 1) + 1)) * out_pitch + (size_t)((oy & 1) * 2 + (ox & 1)) * C + c4 * 4;
    } else if (mode == 4) {
      ooff = ((size_t)(n * (OH + 2) + oy + 1) * (OW + 2) + (ox + 1)) * C + c4 * 4;
    } else {
      ooff = ((size_t)(n * (OH + 2) + oy + 1) * (OW + 2) + (ox + 1)) * out_pitch + c4 * 4;
    }
    *reinterpret_cast<float4*>(out + ooff) = acc;
  }
}

// Tiled variant of modes 1 (up x2) and 2 (down x2): a CTA stages the (affine + activation)-transformed input window of its output
// tile in shared memory once -- the per-tap version above re-evaluates SiLU for every tap, 16x per input element for up x2,
// which made it MUFU-bound at ~3x its HBM time -- then evaluates the 2x2 (up) / 4x4 (down) taps from shared memory.
//   MODE 1: 16x16 outputs <- 10x10 inputs;  MODE 2: 8x8 outputs <- 18x18 inputs;  32 channels per CTA.
template <int MODE>
__global__ void __launch_bounds__(256) fir_pnhwc_tiled_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int act, float* __restrict__ out, int H,
                                                              int W, int C, int out_pitch, float gain, int tiles_x) {
  pdl_wait();
  pdl_trigger();
  constexpr int OT = MODE == 1 ? 16 : 8;             // output tile edge
  constexpr int IT = MODE == 1 ? 10 : 18;            // input window edge
  constexpr int CB = 32, CB4 = CB / 4;
  __shared__ __align__(16) float win[IT * IT * CB];
  const int n = blockIdx.z;
  const int cb0 = blockIdx.y * CB;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int oy0 = ty * OT, ox0 = tx * OT;
  const int OH = MODE == 1 ? 2 * H : H / 2, OW = MODE == 1 ? 2 * W : W / 2;
  const int iy0 = MODE == 1 ? oy0 / 2 - 1 : 2 * oy0 - 1;
  const int ix0 = MODE == 1 ? ox0 / 2 - 1 : 2 * ox0 - 1;
  const bool affine = scale != nullptr;
  const int c4 = threadIdx.x % CB4;
  float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0);
  if (affine) {
    sc = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * C + cb0) + c4);
    sh = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * C + cb0) + c4);
  }
  for (int e = threadIdx.x / CB4; e < IT * IT; e += 256 / CB4) {
    const int wy = e / IT, wx = e - wy * IT;
    const int iy = iy0 + wy, ix = ix0 + wx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);      // outside the image the FIR sees zeros (not act(shift))
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = load_tr(x, ((size_t)(n * (H + 2) + iy + 1) * (W + 2) + (ix + 1)) * C + cb0 + c4 * 4, sc, sh, affine, act);
    reinterpret_cast<float4*>(win)[e * CB4 + c4] = v;
  }
  __syncthreads();
  const float4* w4 = reinterpret_cast<const float4*>(win);
  for (int o = threadIdx.x / CB4; o < OT * OT; o += 256 / CB4) {
    const int ly = o / OT, lx = o - ly * OT;
    const int oy = oy0 + ly, ox = ox0 + lx;
    if (oy >= OH || ox >= OW) continue;
    float4 acc = make_float4(0, 0, 0, 0);
    if (MODE == 1) {
      // even output row: inputs (oy/2 - 1, oy/2) with taps (1, 3); odd: ((oy-1)/2, (oy+1)/2) with taps (3, 1)
      const int wy = (ly >> 1) + (ly & 1), wx = (lx >> 1) + (lx & 1);    // window row of the first contributing input
      const float ay0 = (ly & 1) ? 3.f : 1.f, ay1 = 4.f - ay0;
      const float ax0 = (lx & 1) ? 3.f : 1.f, ax1 = 4.f - ax0;
      const float g = (4.f / 64.f) * gain;
      fma4(acc, w4[((wy) * IT + wx) * CB4 + c4], ay0 * ax0 * g);
      fma4(acc, w4[((wy) * IT + wx + 1) * CB4 + c4], ay0 * ax1 * g);
      fma4(acc, w4[((wy + 1) * IT + wx) * CB4 + c4], ay1 * ax0 * g);
      fma4(acc, w4[((wy + 1) * IT + wx + 1) * CB4 + c4], ay1 * ax1 * g);
    } else {
      const float t4[4] = {1.f, 3.f, 3.f, 1.f};
      const float g = (1.f / 64.f) * gain;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fma4(acc, w4[((2 * ly + i) * IT + 2 * lx + j) * CB4 + c4], t4[i] * t4[j] * g);
    }
    *reinterpret_cast<float4*>(out + ((size_t)(n * (OH + 2) + oy + 1) * (OW + 2) + (ox + 1)) * out_pitch + cb0 + c4 * 4) = acc;
  }
}

}  // namespace ddg

using namespace ddg;

extern "C" int ddg_upfirdn2d_out_size(int in_size, int up, int down, int pad0, int pad1, int ksize) {
  return (in_size * up + pad0 + pad1 - ksize) / down + 1;
}

extern "C" int ddg_upfirdn2d(const float* x, const float* k, float* out, long planes, int in_h, int in_w, int kh, int kw, int up_x,
                             int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, cudaStream_t stream) {
  if (!x || !k || !out || planes < 0 || in_h <= 0 || in_w <= 0 || kh <= 0 || kw <= 0 || up_x < 1 || up_y < 1 || down_x < 1 || down_y < 1) {
    ddg_set_last_error("upfirdn2d: bad args");
    return DDG_ERR_ARG;
  }
  if (kh * kw > kMaxTaps) { ddg_set_last_error("upfirdn2d: kernel larger than 256 taps"); return DDG_ERR_UNSUPPORTED; }
  UpfirdnParams p;
  p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  if (p.out_h <= 0 || p.out_w <= 0) { ddg_set_last_error("upfirdn2d: empty output"); return DDG_ERR_ARG; }
  if (planes == 0) return DDG_OK;
  if (kh == 4 && kw == 4 && up_x == up_y && down_x == down_y && ((up_x == 1 && down_x <= 2) || (up_x == 2 && down_x == 1))) {
    const bool poly = (up_x == 2 && pad_x0 == 2 && pad_y0 == 2 && p.out_h == 2 * in_h && p.out_w == 2 * in_w &&
                       ((((uintptr_t)out) & 7) == 0));
    const long totalf = poly ? planes * in_h * in_w
                             : (up_x == 2 ? planes * p.out_h * ((p.out_w + 3) / 4) : planes * ((p.out_h + 3) / 4) * p.out_w);
    long blk = (totalf + 255) / 256;
    if (blk > 148L * 64) blk = 148L * 64;
    const bool al16 = ((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0;
    if (poly && (in_w % 4 == 0) && al16) {
      long b4 = (planes * in_h * (in_w / 4) + 255) / 256;
      if (b4 > 148L * 64) b4 = 148L * 64;
      upfirdn2d_k4_up2_poly4_kernel<<<(int)b4, 256, 0, stream>>>(x, k, out, planes, in_h, in_w);
    } else if (up_x == 1 && down_x == 2 && pad_x0 == 1 && pad_y0 == 1 && (in_w % 4 == 0) && p.out_w * 2 == in_w && al16) {
      long b4 = (planes * ((p.out_h + 3) / 4) * (p.out_w / 2) + 255) / 256;
      if (b4 > 148L * 64) b4 = 148L * 64;
      upfirdn2d_k4_down2_vec_kernel<<<(int)b4, 256, 0, stream>>>(x, k, out, planes, in_h, in_w, p.out_h, p.out_w);
    } else if (poly)
      upfirdn2d_k4_up2_poly_kernel<<<(int)blk, 256, 0, stream>>>(x, k, out, planes, in_h, in_w);
    else if (up_x == 2)
      upfirdn2d_k4_up2_kernel<<<(int)blk, 256, 0, stream>>>(x, k, out, planes, in_h, in_w, p.out_h, p.out_w, pad_x0, pad_y0);
    else if (down_x == 2)
      upfirdn2d_k4_down_kernel<2><<<(int)blk, 256, 0, stream>>>(x, k, out, planes, in_h, in_w, p.out_h, p.out_w, pad_x0, pad_y0);
    else
      upfirdn2d_k4_down_kernel<1><<<(int)blk, 256, 0, stream>>>(x, k, out, planes, in_h, in_w, p.out_h, p.out_w, pad_x0, pad_y0);
    DDG_CHECK_LAUNCH();
    return DDG_OK;
  }
  const bool vec = (p.out_w % 4 == 0);
  const long total = planes * p.out_h * (vec ? p.out_w / 4 : p.out_w);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 64) blocks = 148L * 64;
  if (vec) upfirdn2d_kernel<4><<<(int)blocks, 256, 0, stream>>>(x, k, out, planes, p);
  else upfirdn2d_kernel<1><<<(int)blocks, 256, 0, stream>>>(x, k, out, planes, p);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_fir_pnhwc(const float* x, const float* scale, const float* shift, int act, float* out, int N, int H, int W, int C,
                             int mode, int out_pitch, float gain, cudaStream_t stream) {
  if (!x || !out || C % 4 != 0 || mode < 1 || mode > 4 || ((scale == nullptr) != (shift == nullptr))) { ddg_set_last_error("fir_pnhwc: bad args"); return DDG_ERR_ARG; }
  if ((mode == 2 || mode == 3 || mode == 4) && ((H | W) & 1)) { ddg_set_last_error("fir_pnhwc: odd size"); return DDG_ERR_UNSUPPORTED; }
  int OH = mode == 1 ? 2 * H : (mode == 2 ? H / 2 : (mode == 3 ? H + 1 : H));
  int OW = mode == 1 ? 2 * W : (mode == 2 ? W / 2 : (mode == 3 ? W + 1 : W));
  if ((mode == 1 || mode == 2) && C % 32 == 0 && N <= 65535) {
    const int OT = mode == 1 ? 16 : 8;
    const int tiles_x = (OW + OT - 1) / OT, tiles_y = (OH + OT - 1) / OT;
    dim3 grid(tiles_x * tiles_y, C / 32, N);
    if (mode == 1) launch_pdl(fir_pnhwc_tiled_kernel<1>, dim3(grid), dim3(256), 0, stream, x, scale, shift, act, out, H, W, C, out_pitch, gain, tiles_x);
    else launch_pdl(fir_pnhwc_tiled_kernel<2>, dim3(grid), dim3(256), 0, stream, x, scale, shift, act, out, H, W, C, out_pitch, gain, tiles_x);
    DDG_CHECK_LAUNCH();
    return DDG_OK;
  }
  const long total = (long)N * OH * OW * (C / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  launch_pdl(fir_pnhwc_kernel, dim3((int)blocks), dim3(256), 0, stream, x, scale, shift, act, out, N, H, W, C, mode, out_pitch, gain);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
