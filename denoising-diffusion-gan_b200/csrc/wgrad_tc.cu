// Weight gradient of the implicit-GEMM convolution on tcgen05 / TMEM (sm_100a).
//
//   dW[co][ci][tap] = sum over output positions q of  dY[q][co] * X[q + off(tap)][ci]
//
// (the wgrad cuDNN computes for nn.Conv2d in the reference's backward, reached from ddgan.py:459,467,477,506).
// GEMM view per tap: D_tap[M = co][N = ci] += dY^T[M][K = pixels] * X_shifted[K][N]: the contraction runs over pixels, and in
// NHWC the contiguous axis is the channel, so both operands are MN-major.  The shared-memory layout is the same
// [chunk of 8 channels][row][8 x bf16] used by the forward kernel; read as an MN-major SWIZZLE_NONE operand it has
// SBO = chunk pitch (next 8 channels) and LBO = 128 B (next 8 pixel rows), and a filter tap is again a start-address offset of
// off(tap) rows.  dY is stored PNHWC with a zero border, so border positions contribute nothing and need no masking.
//
// One CTA owns a (128 output channels) x (128 input channels) x (one tap row, <= 3 taps) block of dW -- 384 fp32 TMEM columns;
// 32 input channels x all taps for narrow inputs -- and a slice of the pixel space (split-K, factor from a cycle model on the
// host); it streams pixel tiles of 64 rows through a 3-stage smem ring (producer warps convert fp32 -> bf16 hi/lo; BF16x3 as in
// the forward) and accumulates in TMEM across all its tiles.  The partial blocks are then staged in shared memory (the operand
// ring is dead by then), summed across a 2-CTA cluster with vector DSMEM loads, re-read in dW order and added to dW with
// red.global.add so that neighbouring lanes share sectors.  UMMA descriptor addresses are masked to the CTA-local offset:
// inside a cluster the shared window of rank r starts at r << 24.
#include <cstdlib>
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

// ---- PTX wrappers (same as conv_tc.cu; kept local so each translation unit is self-contained) ----
__device__ __forceinline__ uint32_t w_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void w_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void w_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void w_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void w_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void w_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t w_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void w_tmem_ld32(uint32_t taddr, float* v) {
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

constexpr int kWProdWarps = 8;
constexpr int kWThreads = (kWProdWarps + 1) * 32;   // + 1 MMA warp (also owns TMEM alloc)
constexpr int MCO = 128;                             // output channels per CTA (GEMM M)

struct WgradDev {
  const float* x;      // PNHWC source as seen by the conv, pitch xpitch
  const float* dy;     // PNHWC output gradient (zero border), pitch dypitch
  float* dw;           // fp32, dw[co*s_co + ci*s_ci + tap*s_tap] += ...
  int xpitch, dypitch;
  int Mtotal;          // rows of the padded linear space N*Hp*Wp
  int Cout, Cin_real;  // bounds for the final scatter
  int dy_c;            // channels physically present in dy (load bound)
  int ntaps;
  int tapoff[9];       // row offset of each tap (dr*Wp + ds)
  int taps_per_group;  // taps handled by one CTA (their accumulators share the 512 TMEM columns)
  int ngroups;
  long s_co, s_ci, s_tap;
  int tiles_per_cta;   // pixel tiles per split
  int n_tiles;         // total pixel tiles
  int xpitch_b;        // X window chunk pitch in bytes
  int splits_pad;      // split slots per (co, ci, tap-group) block, a multiple of the cluster size; blockIdx.z = grp * splits_pad + split
  int cs;              // cluster size along z: the CTAs of one cluster hold partial sums of the same dW block
  float gain;          // multiplies the block before it is added to dW
  long long* prof;     // optional per-role cycle counters of one CTA
  int order;           // dW walk of the final reduction: 0 = taps fastest, then ci, then co (conv layouts); 1 = co fastest (NIN [in][out])
};

// NCI: input channels per CTA (GEMM N per tap); KT: pixel rows per stage (GEMM K per stage); NST: smem stages.
//   <128, 64, 3>: Cin % 128 == 0 -- one tap row (<= 3 taps) per CTA, 128x128x16 MMAs (same operand-read rate as the forward)
//   < 32, 128, 2>: narrow inputs (3-channel image conv, stddev channel) -- all taps per CTA
template <int PREC, int NCI, int KT, int NST>
__global__ void __launch_bounds__(kWThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradDev p) {
  constexpr int NPL = (PREC == 3) ? 2 : 1;
  constexpr int DY_CH = MCO / 8;
  constexpr int X_CH = NCI / 8;
  constexpr int DY_PITCH = (KT + 1) * 16;            // odd row count: conflict-free chunk-strided stores
  constexpr int DY_PLANE = DY_CH * DY_PITCH;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar_base = w_smem_u32(bars);
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (NST + s); };
  const uint32_t accFull = bar_base + 8u * (2 * NST);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * NST + 1));
  const int x_plane = X_CH * p.xpitch_b;
  const int stage_bytes = NPL * (DY_PLANE + x_plane);
  uint8_t* sbase = smem + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * MCO;
  const int ci0 = blockIdx.y * NCI;
  const int grp = blockIdx.z / p.splits_pad;
  const int split = blockIdx.z - grp * p.splits_pad;
  const int t0 = grp * p.taps_per_group;
  const int t1 = min(t0 + p.taps_per_group, p.ntaps);
  int minoff = p.tapoff[t0], maxoff = p.tapoff[t0];
  for (int t = t0 + 1; t < t1; ++t) { minoff = min(minoff, p.tapoff[t]); maxoff = max(maxoff, p.tapoff[t]); }
  const int xrows = KT + (maxoff - minoff);
  const int tile0 = split * p.tiles_per_cta;
  const int tile1 = min(tile0 + p.tiles_per_cta, p.n_tiles);
  const int my_tiles = max(tile1 - tile0, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { w_mbar_init(full(s), kWProdWarps * 32); w_mbar_init(empty(s), 1); }
    w_mbar_init(accFull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWProdWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(w_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWProdWarps) {
    // ============================ producers: fp32 rows -> bf16 hi/lo operand tiles ============================
    const int tid = threadIdx.x;
    constexpr int NPT = kWProdWarps * 32;
    constexpr int UB = 4;                            // independent 32-byte loads in flight per thread
    auto cvt_store = [&](const float4& a, const float4& b, uint8_t* dst_hi, uint8_t* dst_lo) {
      uint4 hi, lo;
      split_bf16x2(a.x, a.y, hi.x, lo.x); split_bf16x2(a.z, a.w, hi.y, lo.y);
      split_bf16x2(b.x, b.y, hi.z, lo.z); split_bf16x2(b.z, b.w, hi.w, lo.w);
      *reinterpret_cast<uint4*>(dst_hi) = hi;
      if (NPL == 2) *reinterpret_cast<uint4*>(dst_lo) = lo;
    };
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == gridDim.z / 2;
    long long w_empty = 0;
    const long long t_p0 = clock64();
    // Batched path (NCI = 128: 4 dY items + <= 5 X items of 32 bytes per thread and tile): every load of a tile is issued before
    // the first conversion, so the L2 / HBM latency is exposed once per tile instead of three times (dY batch, X batch, X tail:
    // ncu attributes a third of the stall samples to the first use of loaded data; the producer ran at 3.9 K cycles per tile
    // against 2.3 K of MMA work).
    constexpr int UDY = (KT * DY_CH) / NPT;          // 4 for KT = 64
    constexpr int UXM = 5;
    const bool batched = (NCI == 128) && ((KT * DY_CH) % NPT == 0) && (xrows * X_CH <= UXM * NPT);
    if (batched) {
      const int nx = xrows * X_CH;
      for (int it = 0; it < my_tiles; ++it) {
        const int st = it % NST;
        const uint32_t ph = (it / NST) & 1;
        const int q0 = (tile0 + it) * KT;
        float4 a[UDY + UXM], b[UDY + UXM];
#pragma unroll
        for (int u = 0; u < UDY; ++u) {
          const int item = tid + u * NPT;
          const int c = item % DY_CH, e = item / DY_CH;
          const int q = q0 + e;
          a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < p.Mtotal && co0 + c * 8 < p.dy_c) {
            const float4* src = reinterpret_cast<const float4*>(p.dy + (size_t)q * p.dypitch + co0 + c * 8);
            a[u] = __ldg(src); b[u] = __ldg(src + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < UXM; ++u) {
          const int item = tid + u * NPT;
          const int c = item % X_CH, e = item / X_CH;
          const int g = q0 + minoff + e;
          a[UDY + u] = b[UDY + u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (item < nx && g >= 0 && g < p.Mtotal) {
            const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)g * p.xpitch + ci0 + c * 8);
            a[UDY + u] = __ldg(src); b[UDY + u] = __ldg(src + 1);
          }
        }
        { const long long tw = clock64(); w_mbar_wait(empty(st), ph ^ 1); w_empty += clock64() - tw; }
        uint8_t* sdy = sbase + st * stage_bytes;
        uint8_t* sx = sdy + NPL * DY_PLANE;
#pragma unroll
        for (int u = 0; u < UDY; ++u) {
          const int item = tid + u * NPT;
          const int c = item % DY_CH, e = item / DY_CH;
          cvt_store(a[u], b[u], sdy + c * DY_PITCH + e * 16, sdy + DY_PLANE + c * DY_PITCH + e * 16);
        }
#pragma unroll
        for (int u = 0; u < UXM; ++u) {
          const int item = tid + u * NPT;
          if (item < nx) {
            const int c = item % X_CH, e = item / X_CH;
            cvt_store(a[UDY + u], b[UDY + u], sx + c * p.xpitch_b + e * 16, sx + x_plane + c * p.xpitch_b + e * 16);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        w_mbar_arrive(full(st));
      }
    } else
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % NST;
      const uint32_t ph = (it / NST) & 1;
      const int q0 = (tile0 + it) * KT;
      { const long long tw = clock64(); w_mbar_wait(empty(st), ph ^ 1); w_empty += clock64() - tw; }
      uint8_t* sdy = sbase + st * stage_bytes;
      uint8_t* sx = sdy + NPL * DY_PLANE;
      // dY tile: KT rows x 16 chunks
      for (int base = tid; base < KT * DY_CH; base += NPT * UB) {
        float4 a[UB], b[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int item = base + u * NPT;
          const int c = item % DY_CH, e = item / DY_CH;
          const int q = q0 + e;
          a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (item < KT * DY_CH && q < p.Mtotal && co0 + c * 8 < p.dy_c) {
            const float4* src = reinterpret_cast<const float4*>(p.dy + (size_t)q * p.dypitch + co0 + c * 8);
            a[u] = __ldg(src); b[u] = __ldg(src + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int item = base + u * NPT;
          if (item < KT * DY_CH) {
            const int c = item % DY_CH, e = item / DY_CH;
            cvt_store(a[u], b[u], sdy + c * DY_PITCH + e * 16, sdy + DY_PLANE + c * DY_PITCH + e * 16);
          }
        }
      }
      // X window: xrows rows x X_CH chunks, rows q0 + minoff ...
      const int nx = xrows * X_CH;
      for (int base = tid; base < nx; base += NPT * UB) {
        float4 a[UB], b[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int item = base + u * NPT;
          const int c = item % X_CH, e = item / X_CH;
          const int g = q0 + minoff + e;
          a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (item < nx && g >= 0 && g < p.Mtotal) {
            const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)g * p.xpitch + ci0 + c * 8);
            a[u] = __ldg(src); b[u] = __ldg(src + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int item = base + u * NPT;
          if (item < nx) {
            const int c = item % X_CH, e = item / X_CH;
            cvt_store(a[u], b[u], sx + c * p.xpitch_b + e * 16, sx + x_plane + c * p.xpitch_b + e * 16);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      w_mbar_arrive(full(st));
    }
    // ============================ epilogue 1/2: TMEM -> shared-memory staging in dW walk order ============================
    // (the operand ring is dead once accFull fires).  Row pitches are chosen so that the 32 lanes (= 32 output channels) of one
    // store hit 32 banks; split slots past the last pixel tile stage zeros.
    const long long t_p1 = clock64();
    w_mbar_wait(accFull, 0);
    const long long t_p2 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float* S = reinterpret_cast<float*>(sbase);
      const int quad = warp & 3, half = warp >> 2;
      const int col = quad * 32 + lane;                // output channel inside the block
      constexpr int NCH = NCI / 32;
      const int ntl = t1 - t0;
      const int njobs = ntl * NCH;
      const int pa0 = NCI * ntl + 4;                   // order 0: S[co][tl][ci], 16-byte aligned rows, pitch = 4 (mod 32) words
      for (int job = half; job < njobs; job += 2) {
        const int tl = job / NCH, ch = job % NCH;
        float v[32];
        w_tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(tl * NCI + ch * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (my_tiles <= 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (p.order == 0) {
          float4* dst = reinterpret_cast<float4*>(S + col * pa0 + tl * NCI + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) S[(tl * NCI + ch * 32 + j) * (MCO + 4) + col] = v[j];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (prof_on && tid == 0) {
      p.prof[0] = t_p1 - t_p0;            // producer loop
      p.prof[1] = w_empty;                // ... waiting for a free stage
      p.prof[2] = t_p2 - t_p1;            // waiting for the last MMAs
      p.prof[3] = clock64() - t_p2;       // TMEM -> smem staging
      p.prof[4] = my_tiles;
    }
  } else {
    // ============================ MMA issuer ============================
    // converged warp, one elected lane issues; descriptors advance with one 32-bit add (see conv_tc.cu)
    {
      uint32_t leader;
      asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
      // MN-major A and B (bits 15, 16), fp32 accumulate, bf16 inputs, N = NCI, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NCI >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // MN-major SWIZZLE_NONE: LBO = 128 B (next 8 K rows), SBO = chunk pitch (next 8 M/N channels)
      const uint32_t lbo_f = ((128u >> 4) & 0x3FFF) << 16;
      const uint32_t a_hi32 = (((uint32_t)DY_PITCH >> 4) & 0x3FFF) | (1u << 14);
      const uint32_t b_hi32 = (((uint32_t)p.xpitch_b >> 4) & 0x3FFF) | (1u << 14);
      long long w_full = 0;
      const long long t_m0 = clock64();
      for (int it = 0; it < my_tiles; ++it) {
        const int st = it % NST;
        { const long long tw = clock64(); w_mbar_wait(full(st), (it / NST) & 1); w_full += clock64() - tw; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) {
          const uint32_t dy_hi = lbo_f | ((w_smem_u32(sbase + st * stage_bytes) >> 4) & 0x3FFFu);   // CTA-local offset: in a cluster the window base is rank << 24
          const uint32_t dy_lo = lbo_f | ((w_smem_u32(sbase + st * stage_bytes + DY_PLANE) >> 4) & 0x3FFFu);   // CTA-local offset: in a cluster the window base is rank << 24
          const uint32_t x_hi = lbo_f | ((w_smem_u32(sbase + st * stage_bytes + NPL * DY_PLANE) >> 4) & 0x3FFFu);   // CTA-local offset: in a cluster the window base is rank << 24
          const uint32_t x_lo = lbo_f | ((w_smem_u32(sbase + st * stage_bytes + NPL * DY_PLANE + x_plane) >> 4) & 0x3FFFu);   // CTA-local offset: in a cluster the window base is rank << 24
          for (int t = t0; t < t1; ++t) {
            const uint32_t xo = (uint32_t)(p.tapoff[t] - minoff);          // rows are 16 B: row offset == 16-byte units
            const uint32_t d = tmem_base + (uint32_t)((t - t0) * NCI);
#pragma unroll
            for (int kk = 0; kk < KT / 16; ++kk) {
              const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
              const uint32_t ko = (uint32_t)kk * 16u;                       // 16 pixel rows = 256 B
              auto mma = [&](uint32_t alo, uint32_t blo, uint32_t accf) {
                asm volatile(
                    "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(a_hi32), "r"(blo), "r"(b_hi32),
                    "r"(idesc), "r"(accf)
                    : "memory");
              };
              if (PREC == 3) {
                mma(dy_lo + ko, x_hi + xo + ko, acc);
                mma(dy_hi + ko, x_lo + xo + ko, 1u);
                mma(dy_hi + ko, x_hi + xo + ko, 1u);
              } else {
                mma(dy_hi + ko, x_hi + xo + ko, acc);
              }
            }
          }
          w_commit(empty(st));
        }
        __syncwarp();
      }
      if (leader) w_commit(accFull);
      if (leader && p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == gridDim.z / 2) {
        p.prof[5] = clock64() - t_m0;       // MMA issue loop
        p.prof[6] = w_full;                 // ... waiting for operands
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == kWProdWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
  // ============================ epilogue 2/2: split-K reduction across the cluster, then one reduction into dW ============================
  // Cluster rank r sums rows r, r + cs, ... of the staged block over all cs CTAs (distributed shared memory, contiguous reads) and
  // issues red.global.add in dW order: cs x fewer L2 reductions than one scatter per CTA, and neighbouring lanes share sectors.
  const int cs = p.cs;
  uint32_t rank = 0;
  if (cs > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp < kWProdWarps) {
    const int ntl = t1 - t0;
    const int rows = p.order == 0 ? MCO : ntl * NCI;
    const int rlen = p.order == 0 ? NCI * ntl : MCO;          // multiple of 128 floats either way
    const int pitch = rlen + 4;
    const uint32_t s_local = w_smem_u32(sbase);
    // per-warp scratch row behind the staging block: the summed row is re-read in dW order for sector-sharing reductions
    float* T = reinterpret_cast<float*>(sbase) + (size_t)rows * pitch + (size_t)warp * rlen;
    for (int a = (int)rank + cs * warp; a < rows; a += cs * kWProdWarps) {
      uint32_t rbase[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        rbase[q] = s_local + (uint32_t)(a * pitch) * 4u;
        if (cs > 1 && q < cs) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbase[q]) : "r"(rbase[q]), "r"((uint32_t)q));
      }
      for (int b4 = lane; b4 < rlen / 4; b4 += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cs > 1) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q < cs) {
              float4 v;
              asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(rbase[q] + (uint32_t)b4 * 16u) : "memory");
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
          }
        } else {
          acc = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sbase) + (size_t)a * pitch)[b4];
        }
        reinterpret_cast<float4*>(T)[b4] = acc;
      }
      __syncwarp();
      if (p.order == 0) {
        const int co = co0 + a;
        if (co < p.Cout) {
          float* dst = p.dw + (size_t)co * p.s_co;
          for (int f = lane; f < rlen; f += 32) {              // f walks (ci, tap) with the tap fastest, as dW does
            const int ci_l = f / ntl, tl = f - ci_l * ntl;
            const int ci = ci0 + ci_l;
            if (ci < p.Cin_real) atomicAdd(dst + (size_t)ci * p.s_ci + (size_t)(t0 + tl) * p.s_tap, p.gain * T[tl * NCI + ci_l]);
          }
        }
      } else {
        const int tl = a / NCI, ci = ci0 + (a - tl * NCI);
        if (ci < p.Cin_real) {
          float* dst = p.dw + (size_t)ci * p.s_ci + (size_t)(t0 + tl) * p.s_tap;
          for (int b = lane; b < rlen; b += 32) {
            const int co = co0 + b;
            if (co < p.Cout) atomicAdd(dst + (size_t)co * p.s_co, p.gain * T[b]);
          }
        }
      }
      __syncwarp();
    }
  }
  // peers may still be reading this CTA's staging buffer
  if (cs > 1) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int PREC, int NCI, int KT, int NST>
static int launch_wgrad(WgradDev& d, int Cout, int Cin_pad, int group_taps, cudaStream_t stream) {
  d.taps_per_group = group_taps;
  d.ngroups = (d.ntaps + group_taps - 1) / group_taps;
  int span = 0;
  for (int g = 0; g < d.ngroups; ++g) {
    int lo = d.tapoff[g * group_taps], hi = lo;
    for (int t = g * group_taps; t < d.ntaps && t < (g + 1) * group_taps; ++t) { lo = d.tapoff[t] < lo ? d.tapoff[t] : lo; hi = d.tapoff[t] > hi ? d.tapoff[t] : hi; }
    if (hi - lo > span) span = hi - lo;
  }
  int rows = KT + span;
  if ((rows & 1) == 0) rows += 1;
  d.xpitch_b = rows * 16;
  d.n_tiles = (d.Mtotal + KT - 1) / KT;
  constexpr int NPL = PREC == 3 ? 2 : 1;
  const size_t stage = (size_t)NPL * ((MCO / 8) * (KT + 1) * 16 + (NCI / 8) * d.xpitch_b);
  size_t smem = 128 + NST * stage;
  {
    // the staging buffer of the final reduction reuses the operand ring
    // rows x (row + 4) floats in either walk order, + one scratch row per reducing warp
    const size_t r0 = (size_t)MCO * ((size_t)NCI * group_taps + 4), r1 = (size_t)NCI * group_taps * (MCO + 4);
    const size_t staging = ((r0 > r1 ? r0 : r1) + (size_t)kWProdWarps * (NCI * group_taps > MCO ? NCI * group_taps : MCO)) * 4;
    if (128 + staging > smem) smem = 128 + staging;
  }
  if (smem > 227 * 1024) { ddg_set_last_error("conv2d_wgrad: shared memory budget exceeded (image too wide)"); return DDG_ERR_UNSUPPORTED; }
  const int gx = (Cout + MCO - 1) / MCO, gy = Cin_pad / NCI;
  const int base = gx * gy * d.ngroups;
  // Split-K factor from a small cost model (cycles).  A cluster of cs CTAs reduces its partial blocks through distributed shared
  // memory and issues one set of 128 x NCI x taps reductions into dW (the L2 retires ~100 scattered fp32 reductions per clock
  // chip-wide; walked in dW order they share sectors, ~3x cheaper), so more splits shorten the serial MMA chain at little cost.
  int splits = 1, cs = 1;
  {
    const double mma_cyc = (NCI >= 128 ? 64.0 : 36.0) * group_taps * (KT / 16) * (PREC == 3 ? 3 : 1);
    const double chunks = (double)(KT * (MCO / 8) + (KT + span) * (NCI / 8)) / (kWProdWarps * 32);
    const double prod_cyc = chunks * 35.0 * 2.0;     // ~35 instructions per 16-byte chunk, 2 producer warps per scheduler
    const double t_tile = mma_cyc > prod_cyc ? mma_cyc : prod_cyc;
    const double block_elems = 128.0 * NCI * group_taps;
    const double t_fixed = 8000.0 + block_elems / 32.0;   // setup + TMEM -> smem staging
    double best = 1e30;
    const int smax = d.n_tiles < 4 * 148 ? d.n_tiles : 4 * 148;
    // measured on B200: CTA pairs (one TPC) always co-schedule; clusters of 4 / 8 full-SM CTAs wait for whole-GPC slots and lose 1.5-2x
    static int cs_max = getenv("DDG_WGRAD_CS_MAX") ? atoi(getenv("DDG_WGRAD_CS_MAX")) : 2;
    for (int c2 = 1; c2 <= cs_max; c2 *= 2) {
      for (int sp = c2; sp <= smax || sp == c2; sp += c2) {
        if (sp > d.n_tiles && sp > c2) break;
        const int tpc = (d.n_tiles + sp - 1) / sp;
        const double ctas = (double)base * sp;
        const double waves = (double)(((long)ctas + 147) / 148);
        const double t_red = c2 > 1 ? 3000.0 + block_elems / 64.0 : 0.0;   // cluster barriers + DSMEM sweep of this CTA's rows
        const double t = waves * (tpc * t_tile + t_fixed + t_red) + (ctas / c2) * block_elems / 300.0;
        if (t < best) { best = t; splits = sp; cs = c2; }
      }
    }
  }
  d.tiles_per_cta = (d.n_tiles + splits - 1) / splits;
  d.splits_pad = splits;
  d.cs = cs;
  d.order = (d.s_co < d.s_ci) ? 1 : 0;
  auto kern = wgrad_tc_kernel<PREC, NCI, KT, NST>;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, gy, d.ngroups * splits);
  cfg.blockDim = dim3(kWThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = cs;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, d);
  if (e != cudaSuccess) { ddg_set_last_error(cudaGetErrorString(e)); return DDG_ERR_LAUNCH; }
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

}  // namespace ddg

using namespace ddg;

extern "C" int ddg_conv2d_wgrad(const ddg_wgrad_desc* c, cudaStream_t stream) {
  if (!c || !c->x || !c->dy || !c->dw || c->ntaps < 1 || c->ntaps > 9) { ddg_set_last_error("conv2d_wgrad: bad args"); return DDG_ERR_ARG; }
  if (c->Cin_pad % 32 != 0 || c->xpitch % 4 != 0 || c->dypitch % 4 != 0 || c->dy_cpad % 8 != 0) {
    ddg_set_last_error("conv2d_wgrad: channel counts must be padded (Cin to 32, pitches to 4)");
    return DDG_ERR_ARG;
  }
  WgradDev d{};
  d.x = c->x; d.dy = c->dy; d.dw = c->dw;
  d.xpitch = c->xpitch; d.dypitch = c->dypitch;
  d.Mtotal = c->N * c->Hp * c->Wp;
  d.Cout = c->Cout;
  d.dy_c = c->dy_cpad;
  d.Cin_real = c->Cin_real;
  d.ntaps = c->ntaps;
  for (int t = 0; t < c->ntaps; ++t) d.tapoff[t] = c->tap_dr[t] * c->Wp + c->tap_ds[t];
  d.s_co = c->s_co; d.s_ci = c->s_ci; d.s_tap = c->s_tap;
  d.gain = c->gain == 0.f ? 1.f : c->gain;
  d.prof = (long long*)c->debug_prof;
  const int prec = c->precision == 1 ? 1 : 3;
  // one tap row per CTA: 3x3 -> 3 groups of 3, 2x2 -> 2 groups of 2, 1x1 -> 1 group
  const int gt = c->ntaps == 9 ? 3 : (c->ntaps == 4 ? 2 : (c->ntaps <= 4 ? c->ntaps : 3));
  if (c->Cin_pad % 128 == 0) {
    return prec == 3 ? launch_wgrad<3, 128, 64, 3>(d, c->Cout, c->Cin_pad, gt, stream)
                     : launch_wgrad<1, 128, 64, 3>(d, c->Cout, c->Cin_pad, gt, stream);
  }
  if (c->Cin_pad % 64 == 0) {
    // 64-channel maps (the 256-px NCSN++ configurations, ch = 64): N = 64 per tap, one tap row per CTA
    return prec == 3 ? launch_wgrad<3, 64, 64, 3>(d, c->Cout, c->Cin_pad, gt, stream)
                     : launch_wgrad<1, 64, 64, 3>(d, c->Cout, c->Cin_pad, gt, stream);
  }
  // narrow inputs (image convs, stddev channel): all taps per CTA while the X window (KT + span rows) fits, else one tap row
  int rc = prec == 3 ? launch_wgrad<3, 32, 128, 2>(d, c->Cout, c->Cin_pad, c->ntaps, stream)
                     : launch_wgrad<1, 32, 128, 2>(d, c->Cout, c->Cin_pad, c->ntaps, stream);
  if (rc == DDG_ERR_UNSUPPORTED && gt < c->ntaps)
    rc = prec == 3 ? launch_wgrad<3, 32, 128, 2>(d, c->Cout, c->Cin_pad, gt, stream)
                   : launch_wgrad<1, 32, 128, 2>(d, c->Cout, c->Cin_pad, gt, stream);
  return rc;
}
