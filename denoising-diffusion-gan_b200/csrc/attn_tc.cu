// Fused attention core of AttnBlockpp (score_sde/models/layerspp.py:108-124) on tcgen05 / TMEM, sm_100a.
//
//   w = softmax(q k^T / sqrt(C))  (einsum :115-118),  h = w v (:119),  out = (x + NIN_3(h)) / sqrt(2)  (:120-124)
//
// for T = H*W = 256 tokens and C = 256 channels (the 16x16 attention level of the CIFAR-10 and the 256-px NCSN++ configurations).
// One CTA owns (sample n, 128 queries) and chains three GEMMs without leaving the SM -- the logits S, the weights P and the
// attention output O never touch HBM, and there are no per-image operand packs:
//
//   phase 1  S[128 x 256 keys]  = Q K^T        A = Q, B = K, both converted fp32 -> bf16 hi/lo by the worker warps, K-blocks of 32 ch
//   phase 2  P~ = exp(S/16 - rowmax)           TMEM -> registers -> bf16 hi/lo planes in shared memory (A operand of phase 3);
//                                              the row sums stay in registers and are divided out of O (flash-attention style)
//   phase 3  O[128 x 256 ch]    = P~ V         B = V read as an MN-major operand (tokens are the contraction axis, channels contiguous)
//   phase 4  Y[128 x 256]       = (O / rowsum) W3     A = O from TMEM re-split into shared memory, B = packed NIN_3 weights (bulk TMA)
//   epilogue out = (Y + b3 + x) * out_scale, per-(n, c) sum / sum-of-squares for the GroupNorm that follows
//
// Every GEMM is BF16x3 (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM) in the fp32-parity mode, single-pass BF16 in precision 1.
// TMEM: S / Y in columns [0, 256), O in [256, 512).  Shared memory: phase 1 runs a 4-stage (Q block | K block) ring; from phase 2 on
// the same bytes hold the P~ / O operand (128 rows x 256, chunked K-major) and a 2-stage ring used first for V blocks, then for the
// NIN_3 weight stages.
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {
namespace attn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

constexpr int T = 256, C = 256;          // tokens per sample, channels
constexpr int MQ = 128;                  // queries per CTA
constexpr int kWorkWarps = 8, kWork = kWorkWarps * 32;
constexpr int kThreads = kWork + 64;     // + loader warp + MMA warp
constexpr int PO_PITCH = (MQ + 2) * 16;  // chunk pitch of the P~ / O operand and of the Q block: = 32 (mod 128) bytes
constexpr int KB_PITCH = (T + 2) * 16;   // chunk pitch of the K block
constexpr int V_ROWS = 32;               // keys per V stage
constexpr int V_PITCH = (V_ROWS + 1) * 16;   // = 16 (mod 128): the 8 chunk-strided stores of a wavefront hit 8 bank groups
constexpr int NST1 = 4, NST2 = 2;

struct Dev {
  const float* qkv; const uint8_t* w3; const float* bias; const float* res; float* out; double* stats;
  int N, H, W;
  float out_scale;
};

template <int PREC>
__global__ void __launch_bounds__(kThreads, 1) attn_kernel(const __grid_constant__ Dev p) {
  constexpr int NPL = PREC == 3 ? 2 : 1;
  constexpr int A1_PLANE = 4 * PO_PITCH, B1_PLANE = 4 * KB_PITCH;
  constexpr int STAGE1 = NPL * (A1_PLANE + B1_PLANE);
  constexpr int PO_PLANE = (C / 8) * PO_PITCH;
  constexpr int PO_BYTES = NPL * PO_PLANE;
  constexpr int V_PLANE = (C / 8) * V_PITCH;
  constexpr int STAGE2 = NPL * V_PLANE;              // >= one packed NIN_3 stage (NPL * 32 * 256 * 2 bytes)
  constexpr int W3_STAGE = NPL * 32 * C * 2;
  static_assert(W3_STAGE <= STAGE2, "weight stage must fit the V ring slot");
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bb = smem_u32(bars);
  auto fullA = [&](int s) { return bb + 8u * s; };
  auto emptyA = [&](int s) { return bb + 8u * (4 + s); };
  auto fullV = [&](int s) { return bb + 8u * (8 + s); };
  auto emptyV = [&](int s) { return bb + 8u * (10 + s); };
  auto fullW = [&](int s) { return bb + 8u * (12 + s); };
  auto emptyW = [&](int s) { return bb + 8u * (14 + s); };
  const uint32_t accS = bb + 8u * 16, pReady = bb + 8u * 17, accO = bb + 8u * 18, oReady = bb + 8u * 19, accY = bb + 8u * 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 21);
  float* xch = reinterpret_cast<float*>(smem + 192);        // [2][128] row exchange between the two column halves (1 KB)
  uint8_t* base = smem + 192 + 1024;
  uint8_t* ring1 = base;                                    // phase 1: NST1 stages of (Q block | K block)
  uint8_t* po = base;                                       // phases 2-4: P~ then O
  uint8_t* ring2 = base + PO_BYTES;                         // phases 3-4: V blocks, then NIN_3 weight stages

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x >> 1, qh = blockIdx.x & 1;
  const float* qkv_n = p.qkv + (size_t)n * T * (3 * C);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST1; ++s) { mbar_init(fullA(s), kWorkWarps); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < NST2; ++s) { mbar_init(fullV(s), kWorkWarps); mbar_init(emptyV(s), 1); mbar_init(fullW(s), 1); mbar_init(emptyW(s), 1); }
    mbar_init(accS, 1); mbar_init(pReady, kWorkWarps); mbar_init(accO, 1); mbar_init(oReady, kWorkWarps); mbar_init(accY, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWorkWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();      // set-up above overlaps the tail of the previous kernel (PDL launch); q/k/v are read below
  pdl_trigger();

  if (warp < kWorkWarps) {
    const int tid = threadIdx.x;
    auto cvt_store = [&](const float4& a, const float4& b, float mul, uint8_t* dst_hi, uint8_t* dst_lo) {
      uint4 hi, lo;
      split_bf16x2(a.x * mul, a.y * mul, hi.x, lo.x); split_bf16x2(a.z * mul, a.w * mul, hi.y, lo.y);
      split_bf16x2(b.x * mul, b.y * mul, hi.z, lo.z); split_bf16x2(b.z * mul, b.w * mul, hi.w, lo.w);
      *reinterpret_cast<uint4*>(dst_hi) = hi;
      if (NPL == 2) *reinterpret_cast<uint4*>(dst_lo) = lo;
    };
    // ======================= phase 1: Q and K blocks (32 channels each) -> bf16 hi/lo operand tiles =======================
    // item = (row r of [128 queries | 256 keys], 16-byte chunk c of the 32-channel block); consecutive threads read consecutive 32 B
    for (int kb = 0; kb < C / 32; ++kb) {
      const int st = kb % NST1;
      const uint32_t ph = (kb / NST1) & 1;
      float4 a[6], b[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int item = tid + kWork * i;
        const int c = item & 3, r = item >> 2;
        const int tok = r < MQ ? qh * MQ + r : r - MQ;
        const int part = r < MQ ? 0 : 1;
        const float4* src = reinterpret_cast<const float4*>(qkv_n + (size_t)tok * (3 * C) + part * C + kb * 32 + c * 8);
        a[i] = __ldg(src); b[i] = __ldg(src + 1);
      }
      mbar_wait(emptyA(st), ph ^ 1);
      uint8_t* sa = ring1 + st * STAGE1;
      uint8_t* sb = sa + NPL * A1_PLANE;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int item = tid + kWork * i;
        const int c = item & 3, r = item >> 2;
        if (r < MQ) cvt_store(a[i], b[i], 1.f, sa + c * PO_PITCH + r * 16, sa + A1_PLANE + c * PO_PITCH + r * 16);
        else cvt_store(a[i], b[i], 1.f, sb + c * KB_PITCH + (r - MQ) * 16, sb + B1_PLANE + c * KB_PITCH + (r - MQ) * 16);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(fullA(st));
    }
    // ======================= phase 2: softmax over the 256 keys of each query row =======================
    const int quad = warp & 3, half = warp >> 2;          // TMEM lane quadrant; column half [half*128, +128)
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    const float sl2e = 0.0625f * 1.4426950408889634f;     // C^-0.5 * log2(e), C = 256
    mbar_wait(accS, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int ck = 0; ck < 4; ++ck) {
      float v[32];
      tmem_ld32(trow + (uint32_t)(half * 128 + ck * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
    }
    xch[half * MQ + row] = mx;
    named_bar_sync(1, kWork);
    mx = fmaxf(xch[row], xch[MQ + row]);
    named_bar_sync(1, kWork);                             // both halves have read the maxima before the sums overwrite them
    float sum = 0.f;
#pragma unroll 1
    for (int ck = 0; ck < 4; ++ck) {
      float v[32];
      tmem_ld32(trow + (uint32_t)(half * 128 + ck * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((v[j] - mx) * sl2e));
        v[j] = e;
        sum += e;
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = (half * 128 + ck * 32) / 8 + q4;
        uint4 hi, lo;
        split_bf16x2(v[8 * q4], v[8 * q4 + 1], hi.x, lo.x); split_bf16x2(v[8 * q4 + 2], v[8 * q4 + 3], hi.y, lo.y);
        split_bf16x2(v[8 * q4 + 4], v[8 * q4 + 5], hi.z, lo.z); split_bf16x2(v[8 * q4 + 6], v[8 * q4 + 7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(po + chunk * PO_PITCH + row * 16) = hi;
        if (NPL == 2) *reinterpret_cast<uint4*>(po + PO_PLANE + chunk * PO_PITCH + row * 16) = lo;
      }
    }
    xch[half * MQ + row] = sum;
    tc_fence_before();
    fence_proxy_async();
    named_bar_sync(1, kWork);
    const float inv = 1.0f / (xch[row] + xch[MQ + row]);
    __syncwarp();
    if (lane == 0) mbar_arrive(pReady);
    // ======================= phase 3: V blocks of 32 keys, stored for an MN-major B operand =======================
    // item = (key k of the block, 16-byte chunk c of the 256 channels); layout [chunk][key][8 channels]
    for (int j = 0; j < T / V_ROWS; ++j) {
      const int st = j % NST2;
      const uint32_t ph = (j / NST2) & 1;
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int item = tid + kWork * i;
        const int c = item & 31, k = item >> 5;
        const float4* src = reinterpret_cast<const float4*>(qkv_n + (size_t)(j * V_ROWS + k) * (3 * C) + 2 * C + c * 8);
        a[i] = __ldg(src); b[i] = __ldg(src + 1);
      }
      mbar_wait(emptyV(st), ph ^ 1);
      uint8_t* sv = ring2 + st * STAGE2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int item = tid + kWork * i;
        const int c = item & 31, k = item >> 5;
        cvt_store(a[i], b[i], 1.f, sv + c * V_PITCH + k * 16, sv + V_PLANE + c * V_PITCH + k * 16);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(fullV(st));
    }
    // ======================= phase 4 operand: O / rowsum -> bf16 hi/lo planes (over the P~ bytes) =======================
    mbar_wait(accO, 0);
    tc_fence_after();
#pragma unroll 1
    for (int ck = 0; ck < 4; ++ck) {
      float v[32];
      tmem_ld32(trow + (uint32_t)(C + half * 128 + ck * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = (half * 128 + ck * 32) / 8 + q4;
        uint4 hi, lo;
        split_bf16x2(v[8 * q4] * inv, v[8 * q4 + 1] * inv, hi.x, lo.x); split_bf16x2(v[8 * q4 + 2] * inv, v[8 * q4 + 3] * inv, hi.y, lo.y);
        split_bf16x2(v[8 * q4 + 4] * inv, v[8 * q4 + 5] * inv, hi.z, lo.z); split_bf16x2(v[8 * q4 + 6] * inv, v[8 * q4 + 7] * inv, hi.w, lo.w);
        *reinterpret_cast<uint4*>(po + chunk * PO_PITCH + row * 16) = hi;
        if (NPL == 2) *reinterpret_cast<uint4*>(po + PO_PLANE + chunk * PO_PITCH + row * 16) = lo;
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(oReady);
    // ======================= epilogue: (Y + b3 + x) * out_scale -> PNHWC, GroupNorm statistics =======================
    mbar_wait(accY, 0);
    tc_fence_after();
    {
      const int tok = qh * MQ + row;
      const int h = tok / p.W, w = tok - h * p.W;
      const size_t obase = ((size_t)(n * (p.H + 2) + h + 1) * (p.W + 2) + (w + 1)) * C;
#pragma unroll 1
      for (int ck = 0; ck < 4; ++ck) {
        const int col0 = half * 128 + ck * 32;
        float v[32];
        tmem_ld32(trow + (uint32_t)col0, v);
        tmem_ld_wait();
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 b = __ldg(b4 + j); v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w; }
        }
        if (p.res) {
          const float4* r4 = reinterpret_cast<const float4*>(p.res + obase + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 r = __ldg(r4 + j); v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w; }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
        float4* o4 = reinterpret_cast<float4*>(p.out + obase + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        if (p.stats) {
          // lanes = rows, registers = channels: transpose-reduce over the 32 lanes, lane L ends with column L's totals
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1[j] = v[j]; s2[j] = v[j] * v[j]; }
#pragma unroll
          for (int sft = 16; sft >= 1; sft >>= 1) {
            const bool up = (lane & sft) != 0;
#pragma unroll
            for (int i = 0; i < sft; ++i) {
              const float a1 = up ? s1[i] : s1[i + sft];
              const float b1 = up ? s1[i + sft] : s1[i];
              s1[i] = b1 + __shfl_xor_sync(0xffffffffu, a1, sft);
              const float a2 = up ? s2[i] : s2[i + sft];
              const float b2 = up ? s2[i + sft] : s2[i];
              s2[i] = b2 + __shfl_xor_sync(0xffffffffu, a2, sft);
            }
          }
          double* dst = p.stats + ((size_t)n * C + col0 + lane) * 2;
          atomicAdd(dst, (double)s1[0]);
          atomicAdd(dst + 1, (double)s2[0]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == kWorkWarps) {
    // ======================= NIN_3 weight loader (bulk TMA into the V ring slots once phase 3 has drained them) =======================
    if (lane == 0) {
      mbar_wait(accO, 0);
      for (int j = 0; j < C / 32; ++j) {
        const int st = j % NST2;
        const uint32_t ph = (j / NST2) & 1;
        mbar_wait(emptyW(st), ph ^ 1);
        mbar_arrive_expect_tx(fullW(st), W3_STAGE);
        tma_bulk_g2s(smem_u32(ring2 + st * STAGE2), p.w3 + (size_t)j * W3_STAGE, W3_STAGE, fullW(st));
      }
    }
    __syncwarp();
  } else {
    // ======================= MMA issuer =======================
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc_kk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_kmn = idesc_kk | (1u << 16);                        // B operand MN-major (phase 3)
    const uint32_t hi_sbo128 = (128u >> 4) | (1u << 14);                          // descriptor bits 32..63: SBO = 128 B, version 1
    const uint32_t a_lbo = (((uint32_t)PO_PITCH >> 4) & 0x3FFF) << 16;            // K-major A: LBO = chunk pitch
    auto addr16 = [](const void* ptr) { return (smem_u32(ptr) >> 4) & 0x3FFFu; };
    auto mma3 = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t a_h32, uint32_t b_hi, uint32_t b_lo, uint32_t b_h32, uint32_t idesc, uint32_t acc) {
      if (PREC == 3) {
        umma(d, a_lo, a_h32, b_hi, b_h32, idesc, acc);
        umma(d, a_hi, a_h32, b_lo, b_h32, idesc, 1u);
        umma(d, a_hi, a_h32, b_hi, b_h32, idesc, 1u);
      } else {
        umma(d, a_hi, a_h32, b_hi, b_h32, idesc, acc);
      }
    };
    // ---- phase 1: S = Q K^T ----
    {
      const uint32_t b_lbo = (((uint32_t)KB_PITCH >> 4) & 0x3FFF) << 16;
      for (int kb = 0; kb < C / 32; ++kb) {
        const int st = kb % NST1;
        mbar_wait(fullA(st), (kb / NST1) & 1);
        tc_fence_after();
        if (leader) {
          const uint8_t* sa = ring1 + st * STAGE1;
          const uint8_t* sb = sa + NPL * A1_PLANE;
          const uint32_t a_hi = a_lbo | addr16(sa), a_lo = a_lbo | addr16(sa + A1_PLANE);
          const uint32_t b_hi = b_lbo | addr16(sb), b_lo = b_lbo | addr16(sb + B1_PLANE);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t ao = (uint32_t)(kk * 2 * PO_PITCH) >> 4, bo = (uint32_t)(kk * 2 * KB_PITCH) >> 4;
            mma3(tmem, a_hi + ao, a_lo + ao, hi_sbo128, b_hi + bo, b_lo + bo, hi_sbo128, idesc_kk, (kb > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(emptyA(st));
        }
        __syncwarp();
      }
      if (leader) umma_commit(accS);
      __syncwarp();
    }
    // ---- phase 3: O = P~ V (B = V block as MN-major: LBO = 128 B (next 8 keys), SBO = chunk pitch (next 8 channels)) ----
    {
      mbar_wait(pReady, 0);
      tc_fence_after();
      const uint32_t b_lbo = ((128u >> 4) & 0x3FFF) << 16;
      const uint32_t b_h32 = (((uint32_t)V_PITCH >> 4) & 0x3FFF) | (1u << 14);
      const uint32_t p_hi = a_lbo | addr16(po), p_lo = a_lbo | addr16(po + PO_PLANE);
      for (int j = 0; j < T / V_ROWS; ++j) {
        const int st = j % NST2;
        mbar_wait(fullV(st), (j / NST2) & 1);
        tc_fence_after();
        if (leader) {
          const uint8_t* sv = ring2 + st * STAGE2;
          const uint32_t b_hi = b_lbo | addr16(sv), b_lo = b_lbo | addr16(sv + V_PLANE);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t ao = (uint32_t)((4 * j + 2 * kk) * PO_PITCH) >> 4;     // keys 32j + 16kk .. +15 = chunks 4j + 2kk, +1
            const uint32_t bo = (uint32_t)(kk * 16 * 16) >> 4;                    // 16 key rows of 16 B
            mma3(tmem + (uint32_t)C, p_hi + ao, p_lo + ao, hi_sbo128, b_hi + bo, b_lo + bo, b_h32, idesc_kmn, (j > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(emptyV(st));
        }
        __syncwarp();
      }
      if (leader) umma_commit(accO);
      __syncwarp();
    }
    // ---- phase 4: Y = (O / rowsum) W3  (packed weights: [4 chunks][256 rows][16 B] per plane and stage) ----
    {
      mbar_wait(oReady, 0);
      tc_fence_after();
      const uint32_t b_lbo = (((uint32_t)(C * 16) >> 4) & 0x3FFF) << 16;
      const uint32_t o_hi = a_lbo | addr16(po), o_lo = a_lbo | addr16(po + PO_PLANE);
      for (int j = 0; j < C / 32; ++j) {
        const int st = j % NST2;
        mbar_wait(fullW(st), (j / NST2) & 1);
        tc_fence_after();
        if (leader) {
          const uint8_t* sw = ring2 + st * STAGE2;
          const uint32_t b_hi = b_lbo | addr16(sw), b_lo = b_lbo | addr16(sw + 32 * C * 2);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t ao = (uint32_t)((4 * j + 2 * kk) * PO_PITCH) >> 4;
            const uint32_t bo = (uint32_t)(kk * 2 * C * 16) >> 4;
            mma3(tmem, o_hi + ao, o_lo + ao, hi_sbo128, b_hi + bo, b_lo + bo, hi_sbo128, idesc_kk, (j > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(emptyW(st));
        }
        __syncwarp();
      }
      if (leader) umma_commit(accY);
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == kWorkWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int PREC>
static size_t smem_bytes() {
  constexpr int NPL = PREC == 3 ? 2 : 1;
  const size_t s1 = (size_t)NST1 * NPL * (4 * PO_PITCH + 4 * KB_PITCH);
  const size_t s2 = (size_t)NPL * (C / 8) * PO_PITCH + (size_t)NST2 * NPL * (C / 8) * V_PITCH;
  return 192 + 1024 + (s1 > s2 ? s1 : s2);
}

}  // namespace attn
}  // namespace ddg

extern "C" int ddg_attention_fwd(const ddg_attn_desc* d, cudaStream_t stream) {
  using namespace ddg::attn;
  if (!d || !d->qkv || !d->w3pack || !d->out || d->N < 1) { ddg_set_last_error("attention_fwd: bad args"); return DDG_ERR_ARG; }
  if (d->H * d->W != T || d->C != C) { ddg_set_last_error("attention_fwd: the fused kernel covers 256 tokens x 256 channels"); return DDG_ERR_UNSUPPORTED; }
  if ((((uintptr_t)d->qkv | (uintptr_t)d->out | (uintptr_t)d->w3pack | (uintptr_t)(d->res ? d->res : d->out)) & 15) != 0) {
    ddg_set_last_error("attention_fwd: pointers must be 16-byte aligned");
    return DDG_ERR_ARG;
  }
  Dev p{};
  p.qkv = d->qkv; p.w3 = (const uint8_t*)d->w3pack; p.bias = d->bias; p.res = d->res; p.out = d->out; p.stats = d->stats;
  p.N = d->N; p.H = d->H; p.W = d->W; p.out_scale = d->out_scale;
  const int prec = d->precision == 1 ? 1 : 3;
  static bool attr3 = false, attr1 = false;
  if (prec == 3) {
    if (!attr3) { cudaFuncSetAttribute(attn_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr3 = true; }
    ddg::launch_pdl(attn_kernel<3>, dim3(2 * d->N), dim3(kThreads), smem_bytes<3>(), stream, p);
  } else {
    if (!attr1) { cudaFuncSetAttribute(attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr1 = true; }
    ddg::launch_pdl(attn_kernel<1>, dim3(2 * d->N), dim3(kThreads), smem_bytes<1>(), stream, p);
  }
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}
