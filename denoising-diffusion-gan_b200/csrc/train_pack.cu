// Launch-count reducers of the training step (ddgan.py:443-518 runs ~5.5 K kernels per iteration when every small piece of
// bookkeeping is its own launch):
//   * ddg_conv_pack_batch : all B-operand (weight) packs of a network -- forward and transposed (dgrad) layouts -- in ONE launch
//     driven by a device-resident item table, instead of one ddg_conv_pack_weights launch per K segment (~400 per step);
//   * ddg_channel_grads   : bias gradient and per-(sample, channel) "Dense_0(temb)" gradient of a conv from its PNHWC output
//     gradient in one pass (was: fp64 statistics kernel + slice + scale + cast + reduce + cast);
//   * ddg_s2d_weights     : the 3x3 stride-2 weights of conv_downsample_2d (up_or_down_sampling.py:149-183) rearranged for the
//     space-to-depth 2x2-tap formulation, and the adjoint rearrangement for their gradient (was: 9 strided copies each way).
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

__global__ void __launch_bounds__(256) pack_batch_kernel(const ddg_pack_item* __restrict__ items, int n_items, long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    // item lookup: chunk_begin is the exclusive prefix sum of the items' 16-byte chunk counts
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].chunk_begin <= i) lo = mid; else hi = mid - 1;
    }
    const ddg_pack_item& it = items[lo];
    const int NT = it.nt, KB = it.kb;
    const int KCH = KB / 8;
    const int nkb = it.cin_pad / KB;
    // Neighbouring threads read neighbouring weights: taps are the innermost axis of conv weights, then whichever of (ci, co)
    // has the smaller stride.  (Row-fastest order made every 4-byte read its own 32-byte sector: 8x read amplification.)
    long r = i - it.chunk_begin;
    const int tap = (int)(r % it.ntaps); r /= it.ntaps;
    int row, ch, kb;
    if (it.s_ci <= it.s_co) {
      ch = (int)(r % KCH); r /= KCH;
      kb = (int)(r % nkb); r /= nkb;
      row = (int)(r % NT); r /= NT;
    } else {
      row = (int)(r % NT); r /= NT;
      ch = (int)(r % KCH); r /= KCH;
      kb = (int)(r % nkb); r /= nkb;
    }
    const int nt = (int)r;
    const int co = nt * NT + row;
    const int tsrc = it.flip_taps ? (it.ntaps - 1 - tap) : tap;
    uint32_t hi4[4], lo4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = 0.f, b = 0.f;
      const int ci = kb * KB + ch * 8 + 2 * j;
      if (co < it.cout) {
        if (ci < it.cin_real) a = it.w[co * it.s_co + ci * it.s_ci + tsrc * it.s_tap];
        if (ci + 1 < it.cin_real) b = it.w[co * it.s_co + (ci + 1) * it.s_ci + tsrc * it.s_tap];
      }
      split_bf16x2(a, b, hi4[j], lo4[j]);
    }
    const long plane_elems = (long)KB * NT;
    const int npl = it.precision == 3 ? 2 : 1;
    const long stage = it.stage_offset + (long)kb * it.ntaps + tap;
    __nv_bfloat16* blob = reinterpret_cast<__nv_bfloat16*>(it.out) + ((long)nt * it.total_stages + stage) * plane_elems * npl;
    *reinterpret_cast<uint4*>(blob + ((long)ch * NT + row) * 8) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
    if (npl == 2) *reinterpret_cast<uint4*>(blob + plane_elems + ((long)ch * NT + row) * 8) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
  }
}

// grid (C / 32, N, S): block = 8 warps; lane = channel, warp w walks pixels p0 + w, p0 + w + 8, ... of sample n
__global__ void __launch_bounds__(256) channel_grads_kernel(const float* __restrict__ dy, float* __restrict__ dav, float* __restrict__ db,
                                                           int H, int W, int C, int cout, float scale, int dav_stride, int pix_per_split) {
  __shared__ float red[8][33];
  const int n = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int wrp = threadIdx.x >> 5;
  const int HW = H * W;
  const int p0 = blockIdx.z * pix_per_split;
  const int p1 = min(p0 + pix_per_split, HW);
  float acc = 0.f;
  for (int p = p0 + wrp; p < p1; p += 8) {
    const int h = p / W, w = p - h * W;
    acc += __ldg(dy + ((size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1)) * C + c);
  }
  red[wrp][threadIdx.x & 31] = acc;
  __syncthreads();
  if (wrp == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    s *= scale;
    if (c < cout) {
      if (dav) {
        if (gridDim.z == 1) dav[(size_t)n * dav_stride + c] = s;
        else atomicAdd(dav + (size_t)n * dav_stride + c, s);
      }
      if (db) atomicAdd(db + c, s);
    }
  }
}

__global__ void s2d_weights_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int cp, int adjoint) {
  // forward: dst = w2 [Cout][2][2][cp][2][2], src = wt [Cout][Cin][3][3];  adjoint: dst = d(wt), src = d(w2)
  if (!adjoint) {
    const long total = (long)Cout * 16 * cp;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      long r = i;
      const int dx = r & 1; r >>= 1;
      const int dyy = r & 1; r >>= 1;
      const int ci = (int)(r % cp); r /= cp;
      const int px = r & 1; r >>= 1;
      const int py = r & 1; r >>= 1;
      const int co = (int)r;
      const int rr = 2 * dyy + py, ss = 2 * dx + px;
      dst[i] = (ci < Cin && rr < 3 && ss < 3) ? src[((long)co * Cin + ci) * 9 + rr * 3 + ss] : 0.f;
    }
  } else {
    const long total = (long)Cout * Cin * 9;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      long r = i;
      const int ss = (int)(r % 3); r /= 3;
      const int rr = (int)(r % 3); r /= 3;
      const int ci = (int)(r % Cin); r /= Cin;
      const int co = (int)r;
      const int py = rr & 1, dyy = rr >> 1, px = ss & 1, dx = ss >> 1;
      dst[i] = src[(((((long)co * 2 + py) * 2 + px) * cp + ci) * 2 + dyy) * 2 + dx];
    }
  }
}

}  // namespace ddg

using namespace ddg;

extern "C" long ddg_conv_pack_chunks(int cout, int cin_pad, int ntaps, int kb, int nt) {
  const int n_tiles = (cout + nt - 1) / nt;
  return (long)n_tiles * (cin_pad / kb) * ntaps * (kb / 8) * nt;
}

extern "C" int ddg_conv_pack_batch(const ddg_pack_item* items_dev, int n_items, long total_chunks, cudaStream_t stream) {
  if (!items_dev || n_items < 1 || total_chunks < 1) { ddg_set_last_error("conv_pack_batch: bad args"); return DDG_ERR_ARG; }
  long blocks = (total_chunks + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  pack_batch_kernel<<<(int)blocks, 256, 0, stream>>>(items_dev, n_items, total_chunks);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_channel_grads_splits(int N, int H, int W, int C) {
  const long blocks = (long)N * (C / 32);
  if (blocks >= 2 * 148 || H * W <= 64) return 1;
  long s = (2 * 148 + blocks - 1) / blocks;
  const long smax = (H * W + 63) / 64;
  return (int)(s < smax ? s : smax);
}

extern "C" int ddg_channel_grads(const float* dy, float* dav, float* db, int N, int H, int W, int C, int cout, float scale,
                                 int dav_stride, cudaStream_t stream) {
  if (!dy || (!dav && !db) || C % 32 != 0 || cout > C) { ddg_set_last_error("channel_grads: bad args (C must be a multiple of 32)"); return DDG_ERR_ARG; }
  const int S = ddg_channel_grads_splits(N, H, W, C);
  const int pps = (H * W + S - 1) / S;
  channel_grads_kernel<<<dim3(C / 32, N, S), 256, 0, stream>>>(dy, dav, db, H, W, C, cout, scale, dav_stride, pps);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_s2d_weights(const float* src, float* dst, int Cout, int Cin, int cp, int adjoint, cudaStream_t stream) {
  if (!src || !dst || Cin > cp) { ddg_set_last_error("s2d_weights: bad args"); return DDG_ERR_ARG; }
  const long total = adjoint ? (long)Cout * Cin * 9 : (long)Cout * 16 * cp;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 8) blocks = 148L * 8;
  s2d_weights_kernel<<<(int)blocks, 256, 0, stream>>>(src, dst, Cout, Cin, cp, adjoint);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

// ---- image output (test_ddgan.py:190-201: to_range_0_1 then torchvision.utils.save_image per image) ------------------------------
// out[n][h][w][c] = (uint8) clamp((x[n][c][h][w] * scale + shift) * 255 + 0.5, 0, 255): the whole batch leaves the device as bytes in
// the layout image encoders want, instead of N float tensors converted one by one on the host.
namespace ddg {
__global__ void __launch_bounds__(256) images_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int N, int C, int H,
                                                          int W, float scale, float shift) {
  const long total = (long)N * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / ((long)H * W);
    const long p = i - n * (long)H * W;
    for (int c = 0; c < C; ++c) {
      float v = (x[(n * C + c) * (long)H * W + p] * scale + shift) * 255.f + 0.5f;
      v = fminf(fmaxf(v, 0.f), 255.f);
      out[i * C + c] = (uint8_t)v;
    }
  }
}
}  // namespace ddg

extern "C" int ddg_images_to_u8(const float* x, uint8_t* out, int N, int C, int H, int W, float scale, float shift, cudaStream_t stream) {
  if (!x || !out || N < 1 || C < 1) { ddg_set_last_error("images_to_u8: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * H * W;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  ddg::images_to_u8_kernel<<<(int)blocks, 256, 0, stream>>>(x, out, N, C, H, W, scale, shift);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
