// Implicit-GEMM convolution on tcgen05 tensor cores with TMEM accumulators (sm_100a).
//
// Replaces the cuDNN convolutions the reference reaches through nn.Conv2d / F.conv2d
// (score_sde/models/layers.py:114-138, dense_layer.py:73-80, up_or_down_sampling.py:56,183,
//  NIN at layers.py:489-512) together with the elementwise work around them
// (AdaGN scale/shift + SiLU at layerspp.py:279,300; "+= Dense_0(temb)" :298-299; the 1x1 skip conv
//  and "(x + h)/sqrt(2)" :305-310; LeakyReLU at discriminator.py:78-82; tanh at ncsnpp...:428).
//
// GEMM view: D[M = pixels][N = Cout] += A[M][K = taps*Cin] * B[K][N].
//   * A (activations) is produced in shared memory by the producer warps: fp32 NHWC rows are loaded from HBM,
//     the fused prologue y = act(scale[n,c]*x + shift[n,c]) is applied, and the result is split into
//     bf16 hi + bf16 lo planes ("BF16x3": Ahi*Bhi + Alo*Bhi + Ahi*Blo with fp32 accumulation reproduces fp32
//     convolution to ~5e-6 relative L2, SURVEY.md section 7 hard part 1).
//     Layout: [chunk of 8 channels][window row][8 x bf16] = the UMMA K-major SWIZZLE_NONE canonical layout with
//     SBO = 128 B (8 rows x 16 B, rows contiguous) and LBO = window pitch.  Because rows are exactly 16 B apart,
//     a filter tap (dr, ds) is just a start-address offset of (dr*Wp + ds)*16 B into the same window: the 3x3
//     convolution runs over the zero-padded linear pixel space [N][H+2][W+2] and border outputs are discarded.
//   * B (weights) is pre-packed on the device (ddg_conv_pack_weights) into per-stage blobs that are already the
//     shared-memory image; the TMA engine (cp.async.bulk, 1-D) streams them through a ring of mbarrier stages.
//   * D lives in TMEM (MSUB accumulators of 128 lanes x NT fp32 columns); one elected thread issues
//     tcgen05.mma.cta_group::1.kind::f16; tcgen05.commit releases the smem stages and signals the epilogue.
//   * Epilogue: tcgen05.ld -> + bias[c] + addvec[n,c] -> (+ residual) * out_scale -> act -> store (padded NHWC /
//     NHWC / NCHW) and per-(n,c) sum / sum-of-squares accumulation for the GroupNorm that consumes the output.
//   * Tiles: 2-D (16*MSUB image rows x 8 columns, halo window (rows+2) x 10, SBO = 160 B) when H % 16 == 0 and W % 8 == 0,
//     else linear over the padded pixel space.  One tile per CTA (10 warps, producers run the epilogue), or -- when there
//     are more tiles than SMs -- the PERSIST variant: 16 warps, every CTA walks tiles b, b + grid, ..., two accumulator
//     sets in TMEM and four dedicated epilogue warps overlap the epilogue of tile i with the mainloop of tile i + 1,
//     setmaxnreg moves registers from the single-thread roles to the producer warpgroups.
//   * Prologue-free K segments that come with pre-split bf16 planes ([plane][N][C/8][H+2][W+2][8], written by the producing conv's
//     epilogue or ddg_split_planes) are fetched by the TMA engine (cp.async.bulk.tensor.4d, one 160-byte box row per halo-window
//     row) straight into the A ring; the producer warps only keep the barrier count for those stages.
//   * Small spatial levels (a few dozen one-tile CTAs): SPLITK instantiations run 2 or 4 CTAs per output tile over disjoint K
//     ranges; partial accumulators meet in a caller workspace (ddg_conv_desc.splitk_ws), the last rank of a tile reduces.
#include <cstdlib>
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include "common.cuh"
#include "ddgan_b200.h"

namespace ddg {

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// 4-D tiled TMA load (tensor map in kernel-parameter space): box -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_tensor_4d_g2s(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Same MMA with the two 64-bit shared-memory descriptors passed as (lo, hi) 32-bit halves: the hi half (SBO, version) and the
// LBO field of the lo half are loop invariants, so advancing an operand is one 32-bit add on the issuing thread.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                               uint32_t accumulate) {
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
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, P;\n"
      "}\n" : "=r"(pred));
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_NONE, version 1 (Blackwell).  Offsets in bytes (multiples of 16).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ---------------------------------------------------------------------------------------------------------
// Kernel parameters (device view)
// ---------------------------------------------------------------------------------------------------------
struct SrcDev {
  const float* x;
  const float* scale;
  const float* shift;
  int C;          // channels contributed by this source
  int pitch;      // row pitch of x in floats
  int ss_stride;  // row pitch of scale / shift in floats
  int act;
  int ntaps;
  int padded;     // gather mode only: source is padded NHWC
  int tapoff[9];  // window mode: row offset of each tap relative to the window start
  int tma;        // 1: this K segment's A operand comes from pre-split bf16 planes through the TMA engine (2-D tiling only)
  int chunk0;     // first 8-channel chunk of the segment inside the planes tensor
};

struct ConvDev {
  SrcDev src[DDG_CONV_MAX_SRC];
  int nsrc;
  const __nv_bfloat16* wpack;  // [n_tile][stage][blob]
  int window;                  // 1: padded-linear window mode, 0: gather (pure 1x1)
  int N, Hp, Wp;               // window mode: padded input space
  int Hout, Wout;              // output image size
  int Mtotal;                  // rows of the M space
  int Cout;                    // real output channels
  int total_stages;            // B stages per n-tile
  int win_rows;                // rows in the A window (incl. margins)
  int win_pitch;               // LBO of A in bytes
  int margin;                  // rows before the first output row in the window
  const float* bias;
  const float* addvec;
  int addvec_stride;
  const float* res;
  float out_scale;
  int out_act;
  float* out;
  int out_mode;                // 0 padded NHWC, 1 NHWC, 2 NCHW
  int out_C;                   // channel pitch of out / res (NHWC modes)
  double* stats;               // [N][Cout][2] or null
  int tile2d;                  // 1: M tile = (16*MSUB image rows) x 8 columns, window = (16*MSUB+2) x 10 halo (see kernel)
  int tiles_x, tiles_y;        // 2-D tiles per image
  int a_sbo;                   // byte distance between 8-row groups of the A operand (128 linear, 160 = 10*16 for 2-D tiles)
  int sub_stride;              // byte distance between the two 128-row sub-tiles inside the A window
  long long* prof;             // optional: per-role cycle counters of CTA (0,0) (bring-up / tuning aid)
  int batch_rows;              // >0: batched GEMM mode (gather only): blockIdx.z = batch, rows per batch
  int tiles_m, n_tiles;        // persistent variant: tile = n_tile * tiles_m + m_tile, CTA b runs tiles b, b + grid, ...
  long w_batch_stride;         // bytes between the packed B operands of consecutive batches
  int nsa;                     // A-operand ring depth (2..4 stages, whatever the shared-memory budget allows)
  int tma_pitch;               // chunk pitch of a TMA-written A plane: window rows * 16 bytes, dense (producer planes are padded)
  void* out_planes;            // optional pre-split copy of the PNHWC output (see ddg_conv_desc.out_planes)
  long out_plane_bytes;
  // split-K (one-tile CTAs on the small spatial levels): ksplit CTAs share an output tile, CTA rank r runs K blocks
  // [kb_split[r], kb_split[r+1]); ranks 0..ksplit-2 park their fp32 partial accumulators in ws_part, the last rank adds them in its epilogue
  int ksplit;
  int kb_split[5];
  float* ws_part;              // [tile][ksplit-1][128*MSUB][NT] fp32
  unsigned* ws_flag;           // [tile] arrival counters (zero between launches)
  alignas(64) CUtensorMap tmap[DDG_CONV_MAX_SRC][2];   // hi / lo plane of every TMA-fed source
};

constexpr int kProdWarps = 8;
constexpr int kThreads = (kProdWarps + 2) * 32;   // one tile per CTA: the producer warps also run the epilogue
// Persistent variant: 4 warpgroups.  WG0-1 = A producers, WG2 = {B loader, MMA issuer, 2 idle warps}, WG3 = epilogue warps
// (warp % 4 = TMEM lane quadrant).  Registers are moved from WG2 to the producers with setmaxnreg.
constexpr int kThreadsPersist = 16 * 32;
constexpr int kEpiWarp0 = 12;
constexpr int kEpiSlotBytes = 12 * 64 * 4;         // per-warp bias / add-vector staging slots of the epilogue
constexpr int kRegsProd = 168, kRegsUtil = 40, kRegsEpi = 128;   // 256*168 + 128*40 + 128*128 = 64512 <= 65536

template <int MSUB, int NT, int KB, int PREC>
struct ConvCfg {
  static constexpr int MT = 128 * MSUB;
  static constexpr int KCH = KB / 8;                              // 16-byte chunks along K per row
  static constexpr int NPL = (PREC == 3) ? 2 : 1;                 // operand planes (hi, lo)
  static constexpr int B_PLANE = KB * NT * 2;                     // bytes of one B plane per stage
  static constexpr int B_STAGE = B_PLANE * NPL;
  static constexpr int NSB = (B_STAGE <= 8192) ? 6 : (B_STAGE <= 16384 ? 4 : 3);
  static constexpr int TMEM_COLS = (MSUB * NT <= 32) ? 32 : (MSUB * NT <= 64 ? 64 : (MSUB * NT <= 128 ? 128 : (MSUB * NT <= 256 ? 256 : 512)));
};

__device__ __forceinline__ void decode_out_row(const ConvDev& p, int m, int m_end, bool& valid, int& n, int& h, int& w) {
  if (p.window) {
    int img = p.Hp * p.Wp;
    n = m / img;
    int r = m - n * img;
    int hp = r / p.Wp;
    int wp = r - hp * p.Wp;
    h = hp - 1;
    w = wp - 1;
    valid = (m < p.Mtotal) && (h >= 0) && (h < p.Hout) && (w >= 0) && (w < p.Wout);
  } else {
    int img = p.Hout * p.Wout;
    n = m / img;
    int r = m - n * img;
    h = r / p.Wout;
    w = r - h * p.Wout;
    valid = m < m_end;
  }
  if (n >= p.N) n = p.N - 1;
}

// PROF: per-role cycle counters (clock64 pairs around every barrier wait) are compiled in only for the tuning builds that
// tools/conv_prof.py asks for through desc.debug_prof; the production instantiations carry none of that.
// SPLITK: separate instantiations for the split-K launches -- with the K-range tests and the partial-sum code compiled into the common
// kernel, every one-tile launch lost 0.7-1.7 us (instruction fetch of the epilogue, dynamic source lookup at the producers' start).
template <int MSUB, int NT, int KB, int PREC, bool PERSIST, bool PROF, bool SPLITK>
__global__ void __launch_bounds__(PERSIST ? kThreadsPersist : kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvDev p) {
#define DDG_CLK() (PROF ? clock64() : 0LL)
  using Cfg = ConvCfg<MSUB, NT, KB, PREC>;
  constexpr int ACC_COLS = MSUB * NT;                                   // TMEM columns of one accumulator set
  constexpr int TM_COLS = PERSIST ? 2 * Cfg::TMEM_COLS : Cfg::TMEM_COLS;  // persistent: two sets, epilogue(i) overlaps mainloop(i+1)
  static_assert(TM_COLS <= 512, "accumulators exceed TMEM");
  constexpr int MT = Cfg::MT;
  constexpr int KCH = Cfg::KCH;
  constexpr int NPL = Cfg::NPL;
  constexpr int NSB = Cfg::NSB;

  extern __shared__ __align__(128) uint8_t smem[];
  // carve: barriers | tmem ptr | B ring | A ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (4 + s); };
  auto fullB = [&](int s) { return bar_base + 8u * (8 + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (8 + NSB + s); };
  auto accFull = [&](int b) { return bar_base + 8u * (8 + 2 * NSB + b); };
  auto accEmpty = [&](int b) { return bar_base + 8u * (10 + 2 * NSB + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (12 + 2 * NSB));
  const int NSA = p.nsa;
  uint8_t* sB = smem + 256;
  const int a_plane = KCH * p.win_pitch;          // bytes of one A plane
  const int a_stage = a_plane * NPL;
  uint8_t* sA = sB + NSB * Cfg::B_STAGE;
  float* epi_slots = reinterpret_cast<float*>(sA + NSA * a_stage);   // 12 x 64 floats: one slot per epilogue-capable warp (see run_epilogue)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Tile loop: the persistent variant walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; otherwise the CTA owns one tile.
  const int tile_first = PERSIST ? (int)blockIdx.x : 0;
  const int tile_stride = PERSIST ? (int)gridDim.x : 1;
  const int tile_end = PERSIST ? p.tiles_m * p.n_tiles : 1;
  int m0 = 0, m_end = p.Mtotal, ntile = 0;
  // 2-D tile geometry: 8 columns wide so that every image row of the tile is exactly one 8-row UMMA core-matrix group; the
  // A operand is then uniformly strided (SBO = 10 entries) inside a (rows+2) x 10 halo window and a tap is still an offset.
  int t2_n = 0, t2_y0 = 0, t2_x0 = 0;
  auto set_tile = [&](int tile) {
    int mx;
    if (PERSIST) { ntile = tile / p.tiles_m; mx = tile - ntile * p.tiles_m; }
    else { mx = blockIdx.x; ntile = blockIdx.y; }
    m0 = (p.batch_rows > 0 ? blockIdx.z * p.batch_rows : 0) + mx * MT;
    m_end = p.batch_rows > 0 ? (blockIdx.z + 1) * p.batch_rows : p.Mtotal;
    if (p.tile2d) {
      const int tx = mx % p.tiles_x;
      const int r_ = mx / p.tiles_x;
      t2_n = r_ / p.tiles_y;
      t2_y0 = (r_ - t2_n * p.tiles_y) * 16 * MSUB;
      t2_x0 = tx * 8;
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(fullA(s), kProdWarps + 1); mbar_init(emptyA(s), 1); }   // 8 producer warps + the loader
    for (int s = 0; s < NSB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(accFull(b), 1); mbar_init(accEmpty(b), PERSIST ? 4 : kProdWarps); }
    fence_barrier_init();
  }
  if (warp == kProdWarps + 1) {
    tmem_alloc<TM_COLS>(smem_u32(tmem_slot));
#ifdef DDG_ENABLE_PDL
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");   // a co-resident CTA (next kernel, PDL) may allocate
#endif
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barriers, TMEM allocation) ran while the previous kernel on the stream was still draining (PDL launch);
  // nothing below may start before its results are complete and visible.
  pdl_wait();

  // number of K blocks
  int nkb_total = 0;
  for (int s = 0; s < p.nsrc; ++s) nkb_total += p.src[s].C / KB;
  // split-K: this CTA's share of the K blocks (the whole range unless ksplit > 1; never in the persistent variant)
  // (the rank is blockIdx.z: the K-range tests of the single-thread roles stay in uniform registers)
  static_assert(!(SPLITK && PERSIST), "split-K is a one-tile-per-CTA mode");
  const int krank = SPLITK ? (int)blockIdx.z : 0;
  const int kb_lo = SPLITK ? p.kb_split[krank] : 0;
  const int kb_hi = SPLITK ? p.kb_split[krank + 1] : nkb_total;
  const int tile_id = SPLITK ? (int)(blockIdx.y * gridDim.x + blockIdx.x) : 0;

  // =================================== epilogue (one tile) ===================================
  // quad: TMEM lane quadrant of the calling warp; the warp handles column chunks half0, half0 + hstep, ...
  // it: tile iteration of this CTA (selects the accumulator set and the barrier parity).
  // wait_acc = false: the caller has already established that the tile's MMAs completed (see the producers' helper call)
  // eslot: 64 floats of shared memory private to the calling warp.  The per-channel bias and the per-(sample, channel) add vector of a
  // 32-column unit are fetched one unit ahead (one channel per lane, in flight during the previous unit -- the first one during the end of
  // the mainloop) and broadcast to the row-per-lane layout through the slot: the eight dependent 128-bit global loads per operand that
  // used to follow every tcgen05.ld exposed an L2 round trip (1-2k cycles under the producers' load) per unit.
  auto run_epilogue = [&](int it, int quad, int half0, int hstep, bool release, bool wait_acc, float* eslot) {
    const int ab = PERSIST ? (it & 1) : 0;
    constexpr int CW = (NT >= 64) ? 32 : 16;         // columns per tcgen05.ld
    constexpr int NCHUNK = NT / CW;
    float pb = 0.f, pa = 0.f;
    auto prefetch = [&](int u) {
      const int c0 = ntile * NT + (u % NCHUNK) * CW;
      pb = 0.f; pa = 0.f;
      if (CW == 32 && c0 + CW <= p.Cout) {
        if (p.bias) pb = __ldg(p.bias + c0 + lane);
        if (p.addvec && p.tile2d) pa = __ldg(p.addvec + (size_t)t2_n * p.addvec_stride + c0 + lane);
      }
    };
    if (CW == 32 && half0 < MSUB * NCHUNK) prefetch(half0);
    if (wait_acc) mbar_wait(accFull(ab), PERSIST ? ((it >> 1) & 1) : 0);
    tc_fence_after();
    // split-K: ranks 0 .. ksplit-2 only park their partial accumulators; the last rank (highest block index of the tile, so that
    // in-order CTA dispatch has put its writers on the machine before it) waits for them and adds them before the usual epilogue.
    const bool sk_writer = SPLITK && krank < p.ksplit - 1;
    const bool sk_reduce = SPLITK && krank == p.ksplit - 1;
    if (sk_reduce) {
      if (lane == 0) {
        const unsigned want = (unsigned)(p.ksplit - 1) * kProdWarps;     // one arrival per epilogue warp of every writer CTA
        while (ld_acquire_gpu(p.ws_flag + tile_id) < want) { }
      }
      __syncwarp();
    }
    // work units = (sub-tile, column chunk) pairs, dealt round-robin to the warps that share this lane quadrant
    int sub_prev = -1;
    bool valid = false; int n = 0, h = 0, w = 0;
    size_t obase = 0;
    int n_first = 0, n_last = 0;
    {
      for (int u = half0; u < MSUB * NCHUNK; u += hstep) {
        const int sub = u / NCHUNK, ck = u - sub * NCHUNK;
        if (sub != sub_prev) {
          sub_prev = sub;
          const int m = m0 + sub * 128 + quad * 32 + lane;
          if (p.tile2d) {
            const int ml = sub * 128 + quad * 32 + lane;
            n = t2_n; h = t2_y0 + (ml >> 3); w = t2_x0 + (ml & 7);
            valid = (h < p.Hout) && (w < p.Wout);
          } else {
            decode_out_row(p, m, m_end, valid, n, h, w);
          }
          obase = 0;
          if (p.out_mode == 0) obase = ((size_t)(n * (p.Hout + 2) + h + 1) * (p.Wout + 2) + (w + 1)) * p.out_C;
          else if (p.out_mode == 1) obase = ((size_t)(n * p.Hout + h) * p.Wout + w) * p.out_C;
          n_first = __shfl_sync(0xffffffffu, n, 0);
          n_last = __shfl_sync(0xffffffffu, n, 31);
        }
        const int col0 = ntile * NT + ck * CW;
        float v[CW];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * ACC_COLS + sub * NT + ck * CW);
        if (CW == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        if (CW == 32) {
          __syncwarp();                                  // the previous unit has been read out of the slot
          eslot[lane] = pb; eslot[32 + lane] = pa;
          __syncwarp();
          if (u + hstep < MSUB * NCHUNK) prefetch(u + hstep);
        }
        tmem_ld_wait();
        if (sk_writer || sk_reduce) {
          const size_t prow = (size_t)(sub * 128 + quad * 32 + lane) * NT + ck * CW;   // thread = accumulator row: 128 contiguous bytes
          if (sk_writer) {
            float4* dst = reinterpret_cast<float4*>(p.ws_part + ((size_t)(tile_id * (p.ksplit - 1) + krank) * MT) * NT + prow);
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) __stcg(dst + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            continue;
          }
#pragma unroll 1
          for (int r = 0; r < p.ksplit - 1; ++r) {
            const float4* src = reinterpret_cast<const float4*>(p.ws_part + ((size_t)(tile_id * (p.ksplit - 1) + r) * MT) * NT + prow);
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 t = __ldcg(src + j);
              v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
          }
        }
        if (col0 < p.Cout) {
          const bool full = (col0 + CW <= p.Cout);
          if (full) {
            // vectorised per-channel bias and per-(sample, channel) add (Dense_0(temb) / dense_t1)
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < CW / 4; ++j) {
                const float4 b = (CW == 32) ? *reinterpret_cast<const float4*>(eslot + 4 * j)
                                            : __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            if (p.addvec) {
              const bool staged = (CW == 32) && p.tile2d;     // one sample per tile: the add vector came through the slot
              const float4* a4 = reinterpret_cast<const float4*>(p.addvec + (size_t)n * p.addvec_stride + col0);
#pragma unroll
              for (int j = 0; j < CW / 4; ++j) {
                const float4 b = staged ? *reinterpret_cast<const float4*>(eslot + 32 + 4 * j) : __ldg(a4 + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              const int cc = col0 + j;
              if (cc < p.Cout) {
                if (p.bias) v[j] += __ldg(p.bias + cc);
                if (p.addvec) v[j] += __ldg(p.addvec + (size_t)n * p.addvec_stride + cc);
              }
            }
          }
          if (valid) {
            if (p.out_mode != 2) {
              if (p.res) {
                const float4* r4 = reinterpret_cast<const float4*>(p.res + obase + col0);
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                  const float4 r = __ldg(r4 + j);
                  v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
                }
              }
              const float osc = p.out_scale;
              if (p.out_act == ACT_TANH) {
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = tanhf(v[j] * osc);
              } else {
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] *= osc;
              }
              float4* o4 = reinterpret_cast<float4*>(p.out + obase + col0);
#pragma unroll
              for (int j = 0; j < CW / 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (p.out_planes != nullptr && full && p.out_mode == 0) {
                // the same values once more as bf16 hi / lo planes [n][chunk][h+1][w+1][8]: consumers that take this tensor without a
                // prologue (1x1 skip convs) fetch them with the TMA engine instead of converting fp32 in their producer warps
                const size_t img = (size_t)(p.Hout + 2) * (p.Wout + 2);
                const size_t pix = (size_t)(h + 1) * (p.Wout + 2) + (w + 1);
                uint8_t* pb = reinterpret_cast<uint8_t*>(p.out_planes);
#pragma unroll
                for (int q = 0; q < CW / 8; ++q) {
                  uint4 hi, lo;
                  split_bf16x2(v[8 * q], v[8 * q + 1], hi.x, lo.x); split_bf16x2(v[8 * q + 2], v[8 * q + 3], hi.y, lo.y);
                  split_bf16x2(v[8 * q + 4], v[8 * q + 5], hi.z, lo.z); split_bf16x2(v[8 * q + 6], v[8 * q + 7], hi.w, lo.w);
                  const size_t off = (((size_t)n * (p.out_C / 8) + (col0 / 8 + q)) * img + pix) * 16;
                  *reinterpret_cast<uint4*>(pb + off) = hi;
                  if (NPL == 2) *reinterpret_cast<uint4*>(pb + p.out_plane_bytes + off) = lo;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                const int cc = col0 + j;
                if (cc < p.Cout) {
                  const size_t oi = ((size_t)(n * p.Cout + cc) * p.Hout + h) * p.Wout + w;
                  float y = v[j];
                  if (p.res) y += __ldg(p.res + oi);
                  y = apply_act(y * p.out_scale, p.out_act);
                  v[j] = y;
                  p.out[oi] = y;
                }
              }
            }
          }
          if (p.stats) {
            // per-(n, channel) sum and sum of squares of the stored values; lanes = rows, registers = channels.
            for (int pass = 0; pass < 2; ++pass) {
              const int n_sel = pass == 0 ? n_first : n_last;
              if (pass == 1 && n_last == n_first) break;
              const bool mine = valid && (n == n_sel);
              float s1[CW], s2[CW];
#pragma unroll
              for (int j = 0; j < CW; ++j) { const float y = mine ? v[j] : 0.f; s1[j] = y; s2[j] = y * y; }
              // transpose-reduce over the 32 lanes: afterwards lane L holds column (L % CW) totals in s1[0]/s2[0]
#pragma unroll
              for (int sft = 16; sft >= 1; sft >>= 1) {
                if (sft < CW) {
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
                } else {
                  // CW == 16 and sft == 16: plain butterfly add of all 16 columns
#pragma unroll
                  for (int i = 0; i < CW; ++i) {
                    s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], sft);
                    s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], sft);
                  }
                }
              }
              const int cc = col0 + (lane % CW);
              if (cc < p.Cout && (CW == 32 || lane < 16)) {
                double* dst = p.stats + ((size_t)n_sel * p.Cout + cc) * 2;
                atomicAdd(dst, (double)s1[0]);
                atomicAdd(dst + 1, (double)s2[0]);
              }
            }
          }
        }
      }
    }
    if (sk_writer) {                                   // partials of this warp are out: publish them (fence, then one arrival per warp)
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(p.ws_flag + tile_id, 1u);
    }
    // accumulator set drained: hand it back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (release && lane == 0) mbar_arrive(accEmpty(ab));
  };

  if (warp < kProdWarps) {
    // =================================== A producers ===================================
    // Software-pipelined: each thread owns one 16-byte channel chunk (c) and up to IMAX window rows (e0 + k*ESTEP).
    // Row geometry (global offset, sample index, liveness) is computed once per source; the rows of K-block i+1 are
    // prefetched into registers while K-block i is converted and stored, so one global-load latency is exposed per
    // K-block at most (it was one per row before: the dominant long-scoreboard stall in the first ncu capture).
    const int tid = threadIdx.x;                     // 0..255
    constexpr int NPT = kProdWarps * 32;
    constexpr int ESTEP = NPT / KCH;
    constexpr int IMAX = (MSUB == 2) ? 5 : 4;   // 320 threads cap ptxas at 168 registers: a 6th pipelined row would spill
    const int c = tid % KCH;                         // fixed 16-byte chunk of this thread
    const int e0 = tid / KCH;

    // off >= 0: element offset of the row in the source; -1: store zeros; -2: row not part of this source's window
    auto row_info = [&](const SrcDev& S, int e, int row_hi, int& off, int& nn) {
      off = -2; nn = 0;
      if (e >= row_hi) return;
      off = -1;
      if (p.tile2d) {
        const int wy = e / 10, wx = e - wy * 10;
        const int y = t2_y0 - 1 + wy, x = t2_x0 - 1 + wx;     // image coordinates; the PNHWC buffer holds the zero border
        nn = t2_n;
        const bool inside = (y >= 0) && (y < p.Hout) && (x >= 0) && (x < p.Wout);
        if (y <= p.Hout && (inside || S.scale == nullptr))
          off = ((t2_n * (p.Hout + 2) + y + 1) * (p.Wout + 2) + (x + 1)) * S.pitch;
      } else if (p.window) {
        const int g = m0 - p.margin + e;             // row in padded linear space
        if (g >= 0 && g < p.Mtotal) {
          const int img = p.Hp * p.Wp;
          const int n = g / img;
          bool live = true;
          if (S.scale != nullptr) {                  // padding must stay zero after the affine prologue
            const int r = g - n * img;
            const int hp = r / p.Wp, wp = r - hp * p.Wp;
            live = (hp >= 1) && (hp < p.Hp - 1) && (wp >= 1) && (wp < p.Wp - 1);
          }
          if (live) { off = g * S.pitch; nn = n; }
        }
      } else {
        const int m = m0 + e;
        if (m < m_end) {
          const int img = p.Hout * p.Wout;
          const int n = m / img;
          nn = n;
          if (S.padded) {
            const int r = m - n * img;
            const int h = r / p.Wout, w = r - h * p.Wout;
            off = ((n * (p.Hout + 2) + h + 1) * (p.Wout + 2) + (w + 1)) * S.pitch;
          } else {
            off = m * S.pitch;
          }
        }
      }
    };
    auto src_rows = [&](const SrcDev& S, int& row_lo, int& row_hi) {
      row_lo = 0; row_hi = p.win_rows;
      if (p.window && !p.tile2d && S.ntaps == 1) { row_lo = S.tapoff[0]; row_hi = S.tapoff[0] + MT; }
    };
    auto transform_store = [&](const SrcDev& S, float* v, bool live, int n, int ch0, int& cur_n, float* sc, float* sh,
                               uint8_t* dst_hi, uint8_t* dst_lo, int e) {
      if (live) {
        if (S.scale != nullptr) {
          if (n != cur_n) {
            cur_n = n;
            const float4* ps = reinterpret_cast<const float4*>(S.scale + (size_t)n * S.ss_stride + ch0);
            const float4* pt = reinterpret_cast<const float4*>(S.shift + (size_t)n * S.ss_stride + ch0);
            const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), t0 = __ldg(pt), t1 = __ldg(pt + 1);
            sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
            sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
        }
        if (S.act == ACT_SILU) {
          // x * sigmoid(x) as x * rcp(1 + 2^(-x log2 e)): ex2.approx.ftz has no denormal fix-up sequence (3 fewer instructions
          // per element than __expf; the prologue is issue- and MUFU-bound on 1-tap K segments and the small levels)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[j] * -1.4426950408889634f));
            v[j] = __fdividef(v[j], 1.0f + e);
          }
        } else if (S.act == ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = leaky_f(v[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      uint4 hi, lo;
      split_bf16x2(v[0], v[1], hi.x, lo.x);
      split_bf16x2(v[2], v[3], hi.y, lo.y);
      split_bf16x2(v[4], v[5], hi.z, lo.z);
      split_bf16x2(v[6], v[7], hi.w, lo.w);
      *reinterpret_cast<uint4*>(dst_hi + e * 16) = hi;
      if (NPL == 2) *reinterpret_cast<uint4*>(dst_lo + e * 16) = lo;
    };

    if (PERSIST) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsProd));
    int stA_p = 0; uint32_t phA_p = 0;                  // A ring position / parity (carried across tiles)
    int last_st = 0; uint32_t last_ph = 0;              // stage / parity of the last K-block this thread produced
    int it = 0;
    bool prof_on = false;
    long long t_prod0 = 0, w_emptyA = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++it) {
      set_tile(tile);
      prof_on = PROF && (p.prof != nullptr) && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
      if (it == 0) t_prod0 = DDG_CLK();
    float cur[IMAX][8], nxt[IMAX][8];
    int off_c[IMAX], nn_c[IMAX], off_n[IMAX], nn_n[IMAX];
    int s_cur = 0, kb_cur = kb_lo;                     // (source, K block inside it) of the first K block of this CTA
    if (SPLITK) { while (kb_cur >= p.src[s_cur].C / KB) { kb_cur -= p.src[s_cur].C / KB; ++s_cur; } }
    if (!p.src[s_cur].tma) {
      int lo_, hi_;
      src_rows(p.src[s_cur], lo_, hi_);
#pragma unroll
      for (int k = 0; k < IMAX; ++k) {
        row_info(p.src[s_cur], lo_ + e0 + k * ESTEP, hi_, off_c[k], nn_c[k]);
        if (off_c[k] >= 0) {
          const float4* q = reinterpret_cast<const float4*>(p.src[s_cur].x + off_c[k] + kb_cur * KB + c * 8);
          const float4 a = __ldg(q), b = __ldg(q + 1);
          cur[k][0] = a.x; cur[k][1] = a.y; cur[k][2] = a.z; cur[k][3] = a.w;
          cur[k][4] = b.x; cur[k][5] = b.y; cur[k][6] = b.z; cur[k][7] = b.w;
        }
      }
    }
    for (int kb_idx = kb_lo; kb_idx < kb_hi; ++kb_idx) {
      const SrcDev& S = p.src[s_cur];
      // ---- scale / shift of this K-block's first live row, issued ahead of the prefetch so that its latency overlaps the
      //      prefetch issue and the wait for a free A stage (it used to be a dependent load in front of the first FFMA) ----
      const int ch0 = kb_cur * KB + c * 8;
      int cur_n = -1;
      float sc[8], sh[8];
      if (S.scale != nullptr) {
        int nf = -1;
#pragma unroll
        for (int k = IMAX - 1; k >= 0; --k) if (off_c[k] >= 0) nf = nn_c[k];
        if (nf >= 0) {
          cur_n = nf;
          const float4* ps = reinterpret_cast<const float4*>(S.scale + (size_t)nf * S.ss_stride + ch0);
          const float4* pt = reinterpret_cast<const float4*>(S.shift + (size_t)nf * S.ss_stride + ch0);
          const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), t0 = __ldg(pt), t1 = __ldg(pt + 1);
          sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
          sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
        }
      }
      // ---- prefetch the next K-block ----
      int s_nxt = s_cur, kb_nxt = kb_cur + 1;
      if (kb_nxt >= S.C / KB) { s_nxt = s_cur + 1; kb_nxt = 0; }
      const bool has_next = kb_idx + 1 < kb_hi;
      if (has_next && !p.src[s_nxt].tma) {
        const SrcDev& Sn = p.src[s_nxt];
        if (s_nxt != s_cur) {
          int lo_, hi_;
          src_rows(Sn, lo_, hi_);
#pragma unroll
          for (int k = 0; k < IMAX; ++k) row_info(Sn, lo_ + e0 + k * ESTEP, hi_, off_n[k], nn_n[k]);
        } else {
#pragma unroll
          for (int k = 0; k < IMAX; ++k) { off_n[k] = off_c[k]; nn_n[k] = nn_c[k]; }
        }
        const int chn = kb_nxt * KB + c * 8;
#pragma unroll
        for (int k = 0; k < IMAX; ++k) {
          if (off_n[k] >= 0) {
            const float4* q = reinterpret_cast<const float4*>(Sn.x + off_n[k] + chn);
            const float4 a = __ldg(q), b = __ldg(q + 1);
            nxt[k][0] = a.x; nxt[k][1] = a.y; nxt[k][2] = a.z; nxt[k][3] = a.w;
            nxt[k][4] = b.x; nxt[k][5] = b.y; nxt[k][6] = b.z; nxt[k][7] = b.w;
          }
        }
      }
      // ---- convert + store the current K-block ----
      const int st = stA_p;
      const uint32_t ph = phA_p;
      last_st = st; last_ph = ph;
      if (++stA_p == NSA) { stA_p = 0; phA_p ^= 1u; }
      { const long long tw = DDG_CLK(); mbar_wait(emptyA(st), ph ^ 1); w_emptyA += DDG_CLK() - tw; }
      uint8_t* dst_hi = sA + st * a_stage + c * p.win_pitch;
      uint8_t* dst_lo = dst_hi + a_plane;
      int row_lo, row_hi;
      src_rows(S, row_lo, row_hi);
      if (S.tma) row_hi = row_lo;       // TMA-fed K segment: the loader thread fills this stage; the producers only keep the barrier count
#pragma unroll
      for (int k = 0; k < IMAX; ++k) {
        if (!S.tma && off_c[k] != -2) transform_store(S, cur[k], off_c[k] >= 0, nn_c[k], ch0, cur_n, sc, sh, dst_hi, dst_lo, row_lo + e0 + k * ESTEP);
      }
      // rows beyond the register pipeline (very wide windows only)
      for (int e = row_lo + e0 + IMAX * ESTEP; e < row_hi; e += ESTEP) {
        int off, nn;
        row_info(S, e, row_hi, off, nn);
        float v[8];
        if (off >= 0) {
          const float4* q = reinterpret_cast<const float4*>(S.x + off + ch0);
          const float4 a = __ldg(q), b = __ldg(q + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        transform_store(S, v, off >= 0, nn, ch0, cur_n, sc, sh, dst_hi, dst_lo, e);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(fullA(st));       // one arrival per warp: 256 serialized arrivals cost more than the copy
      // ---- rotate the register pipeline ----
      if (has_next) {
#pragma unroll
        for (int k = 0; k < IMAX; ++k) {
          off_c[k] = off_n[k]; nn_c[k] = nn_n[k];
#pragma unroll
          for (int j = 0; j < 8; ++j) cur[k][j] = nxt[k][j];
        }
        s_cur = s_nxt; kb_cur = kb_nxt;
      }
    }

    }
    // this CTA has nothing left to produce: once every CTA of the grid is here (or gone) the next kernel's CTAs may be scheduled and
    // run their set-up on SMs that have drained
    pdl_trigger();
    if (!PERSIST) {
      const long long t_prod1 = DDG_CLK();
      mbar_wait(accFull(0), 0);
      const long long t_epi0 = DDG_CLK();
      run_epilogue(0, warp & 3, warp >> 2, 2, true, true, epi_slots + warp * 64);
      if (prof_on && tid == 0) {
        p.prof[0] = t_prod1 - t_prod0;   // producer loop total
        p.prof[1] = w_emptyA;            // ... of which waiting for a free A stage
        p.prof[2] = t_epi0 - t_prod1;    // waiting for the accumulator after the last A stage
        p.prof[3] = DDG_CLK() - t_epi0;  // epilogue
      }
    }
    // last tile of a persistent CTA: nothing is left to produce, so the producer warps take two thirds of its epilogue
    // (outside the tile loop: the register pipeline of the producer loop is dead here)
    if (PERSIST && prof_on && tid == 0) {
      p.prof[0] = DDG_CLK() - t_prod0;   // producer loops of all tiles of this CTA
      p.prof[1] = w_emptyA;              // ... of which waiting for a free A stage
    }
    // The producers may run several tiles ahead of the MMA issuer (A ring depth), so they must NOT wait on accFull by parity: a
    // parity wait is only sound for a waiter at most one phase behind, and two phases early it is satisfied by the previous use of
    // the same accumulator set (seen as 1e-3 errors on short-K layers with a 4-stage ring).  The empty barrier of the last A stage
    // they produced is a barrier whose phase they do track: tcgen05.commit fires it when every MMA issued so far -- i.e. the whole
    // last tile -- has completed, which is exactly the condition for reading the accumulators.
    if (PERSIST && it > 0) {
      mbar_wait(emptyA(last_st), last_ph);
      run_epilogue(it - 1, warp & 3, 1 + (warp >> 2), 3, false, false, epi_slots + warp * 64);
    }
  } else if (warp < kEpiWarp0) {
    // WG2: the two single-thread roles (+ two idle warps in the persistent layout); their registers go to the producers
    if (PERSIST) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsUtil));
    if (warp == kProdWarps) {
    // =================================== weight loader (TMA bulk) ===================================
    if (lane == 0) {
      long long w_emptyB = 0;
      const long long t_l0 = DDG_CLK();
      uint32_t gb = 0;                                 // B stages issued so far (ring position / parity across tiles)
      int stA_l = 0; uint32_t phA_l = 0;               // A ring position / parity: one arrival per K block (with the TMA bytes when TMA-fed)
      const uint32_t tma_plane = (uint32_t)(KCH * p.tma_pitch);
      for (int tile = tile_first; tile < tile_end; tile += tile_stride) {
        set_tile(tile);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)ntile * p.total_stages * Cfg::B_STAGE +
                              (p.batch_rows > 0 ? (size_t)blockIdx.z * p.w_batch_stride : 0);
        int i = 0, kbi = 0;
        for (int s = 0; s < p.nsrc; ++s) {
          const SrcDev& S = p.src[s];
          for (int kb = 0; kb < S.C / KB; ++kb, ++kbi) {
            if (SPLITK && (kbi < kb_lo || kbi >= kb_hi)) { i += S.ntaps; continue; }     // another CTA's share of K (split-K)
            mbar_wait(emptyA(stA_l), phA_l ^ 1);
            if (S.tma) {
              // halo window (16*MSUB+2 rows x 10 columns) x 4 chunks of 8 channels, from the padded planes [n][chunk][y][x][8]
              const uint32_t dst = smem_u32(sA + stA_l * a_stage);
              mbar_arrive_expect_tx(fullA(stA_l), NPL * tma_plane);
              tma_tensor_4d_g2s(dst, &p.tmap[s][0], t2_x0 * 8, t2_y0, S.chunk0 + kb * KCH, t2_n, fullA(stA_l));
              if (NPL == 2) tma_tensor_4d_g2s(dst + (uint32_t)a_plane, &p.tmap[s][1], t2_x0 * 8, t2_y0, S.chunk0 + kb * KCH, t2_n, fullA(stA_l));
            } else {
              mbar_arrive(fullA(stA_l));
            }
            if (++stA_l == NSA) { stA_l = 0; phA_l ^= 1u; }
            for (int t = 0; t < S.ntaps; ++t, ++i, ++gb) {
              const int st = gb % NSB;
              const uint32_t ph = (gb / NSB) & 1;
              { const long long tw = DDG_CLK(); mbar_wait(emptyB(st), ph ^ 1); w_emptyB += DDG_CLK() - tw; }
              mbar_arrive_expect_tx(fullB(st), Cfg::B_STAGE);
              tma_bulk_g2s(smem_u32(sB + st * Cfg::B_STAGE), wsrc + (size_t)i * Cfg::B_STAGE, Cfg::B_STAGE, fullB(st));
            }
          }
        }
      }
      if (PROF && p.prof != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) {
        p.prof[4] = DDG_CLK() - t_l0;  // loader loop total
        p.prof[5] = w_emptyB;          // ... waiting for a free B stage
      }
    }
    __syncwarp();
    } else if (warp == kProdWarps + 1) {
    // =================================== MMA issuer ===================================
    // The whole warp stays converged (all lanes wait on the barriers); one elected lane issues.  The first version ran this loop
    // inside `if (lane == 0)`, rebuilt four 64-bit descriptors per MMA triple and paid compiler-inserted elect / R2UR sequences
    // per instruction: ~110 cycles per tcgen05.mma *issue*, independent of N -- i.e. issue-bound below N = 256.
    {
      const uint32_t leader = elect_one();
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // K-major SWIZZLE_NONE: LBO = byte distance between the two 8-element K halves of one MMA (chunk pitch),
      // SBO = byte distance between 8-row groups (128 B linear tiles, 160 B for 2-D tiles).  Verified on B200.
      const uint32_t a_hi32 = ((uint32_t)p.a_sbo >> 4) | (1u << 14);                 // descriptor bits 32..63
      const uint32_t b_hi32 = (128u >> 4) | (1u << 14);
      const uint32_t b_lbo_f = (((uint32_t)(NT * 16) >> 4) & 0x3FFF) << 16;
      constexpr uint32_t b_kk16 = (2u * NT * 16u) >> 4;
      const uint32_t a_sub16 = (uint32_t)p.sub_stride >> 4;
      uint32_t bi = 0;                                 // B stages consumed so far (ring position across tiles)
      int stA_m = 0; uint32_t phA_m = 0;               // A ring position / parity
      long long w_fullA = 0, w_fullB = 0, w_accE = 0;
      const long long t_m0 = DDG_CLK();
      int it = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++it) {
      const int ab = PERSIST ? (it & 1) : 0;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(ab * ACC_COLS);
      // the epilogue of the tile that used this accumulator set two iterations ago must have drained it
      { const long long tw = DDG_CLK(); mbar_wait(accEmpty(ab), (PERSIST ? ((it >> 1) & 1) : 0) ^ 1); w_accE += DDG_CLK() - tw; }
      tc_fence_after();
      uint32_t acc = 0;
      int s_m = 0, kb_m = 0;
      for (int kb_idx = 0; kb_idx < nkb_total; ++kb_idx) {
        const SrcDev& S = p.src[s_m];
        if (SPLITK && (kb_idx < kb_lo || kb_idx >= kb_hi)) {        // another CTA's share of K (split-K)
          if (++kb_m >= S.C / KB) { kb_m = 0; ++s_m; }
          continue;
        }
        const int stA = stA_m;
        { const long long tw = DDG_CLK(); mbar_wait(fullA(stA), phA_m); w_fullA += DDG_CLK() - tw; }
        if (++stA_m == NSA) { stA_m = 0; phA_m ^= 1u; }
        // TMA-written stages are dense ([chunk][window row][16 B], pitch = rows * 16); producer-written ones carry the bank padding
        const uint32_t pitch_s = S.tma ? (uint32_t)p.tma_pitch : (uint32_t)p.win_pitch;
        const uint32_t a_lbo_s = ((pitch_s >> 4) & 0x3FFF) << 16;
        const uint32_t a_kk16_s = (2u * pitch_s) >> 4;
        const uint32_t a_hi_lo = a_lbo_s | ((smem_u32(sA + stA * a_stage) >> 4) & 0x3FFFu);
        const uint32_t a_lo_lo = a_lbo_s | ((smem_u32(sA + stA * a_stage + a_plane) >> 4) & 0x3FFFu);
        for (int t = 0; t < S.ntaps; ++t, ++bi) {
          const int stB = bi % NSB;
          { const long long tw = DDG_CLK(); mbar_wait(fullB(stB), (bi / NSB) & 1); w_fullB += DDG_CLK() - tw; }
          tc_fence_after();
          if (leader) {
            const uint32_t b_hi_lo = b_lbo_f | ((smem_u32(sB + stB * Cfg::B_STAGE) >> 4) & 0x3FFFu);
            const uint32_t b_lo_lo = b_lbo_f | ((smem_u32(sB + stB * Cfg::B_STAGE + Cfg::B_PLANE) >> 4) & 0x3FFFu);
            const uint32_t toff16 = (uint32_t)(p.window ? S.tapoff[t] : 0);          // rows are 16 B: row offset == 16-byte units
#pragma unroll
            for (int sub = 0; sub < MSUB; ++sub) {
              const uint32_t d = tmem_acc + (uint32_t)(sub * NT);
              uint32_t acc_s = acc;
#pragma unroll
              for (int kk = 0; kk < KB / 16; ++kk) {
                const uint32_t ao = toff16 + (uint32_t)sub * a_sub16 + (uint32_t)kk * a_kk16_s;
                const uint32_t bo = (uint32_t)kk * b_kk16;
                if (PREC == 3) {
                  umma_bf16_lohi(d, a_lo_lo + ao, a_hi32, b_hi_lo + bo, b_hi32, idesc, acc_s);
                  umma_bf16_lohi(d, a_hi_lo + ao, a_hi32, b_lo_lo + bo, b_hi32, idesc, 1u);
                  umma_bf16_lohi(d, a_hi_lo + ao, a_hi32, b_hi_lo + bo, b_hi32, idesc, 1u);
                } else {
                  umma_bf16_lohi(d, a_hi_lo + ao, a_hi32, b_hi_lo + bo, b_hi32, idesc, acc_s);
                }
                acc_s = 1u;
              }
            }
            umma_commit(emptyB(stB));
          }
          acc = 1u;
          __syncwarp();
        }
        if (leader) umma_commit(emptyA(stA));
        __syncwarp();
        if (++kb_m >= S.C / KB) { kb_m = 0; ++s_m; }
      }
      if (leader) umma_commit(accFull(ab));
      __syncwarp();
      }
      if (leader) {
        if (PROF && p.prof != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0) {
          p.prof[6] = DDG_CLK() - t_m0;  // MMA issue loop total
          p.prof[7] = w_fullA;           // ... waiting for A
          p.prof[8] = w_fullB;           // ... waiting for B
          p.prof[9] = w_accE;            // ... waiting for the epilogue to drain an accumulator set (persistent)
          p.prof[10] = it;               // tiles run by this CTA
        }
      }
      __syncwarp();
    }
    }
  } else if (PERSIST) {
    // =================================== epilogue warps (persistent variant) ===================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    int it = 0;
    long long w_accF = 0;
    const long long t_e0 = DDG_CLK();
    for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++it) {
      set_tile(tile);
      const bool last = tile + tile_stride >= tile_end;
      { const long long tw = DDG_CLK(); mbar_wait(accFull(it & 1), (it >> 1) & 1); w_accF += DDG_CLK() - tw; }
      run_epilogue(it, warp & 3, 0, last ? 3 : 1, true, true, epi_slots + (8 + warp - kEpiWarp0) * 64);
    }
    if (PROF && p.prof != nullptr && blockIdx.x == gridDim.x / 2 && warp == kEpiWarp0 && lane == 0) {
      p.prof[11] = DDG_CLK() - t_e0;   // epilogue warps: whole tile loop
      p.prof[12] = w_accF;             // ... of which waiting for a finished accumulator set
    }
  }

  __syncthreads();
  if (SPLITK && krank == p.ksplit - 1 && threadIdx.x == 0) p.ws_flag[tile_id] = 0u;   // ready for the next launch
  if (warp == kProdWarps + 1) {
    tc_fence_after();
    tmem_dealloc<TM_COLS>(tmem_base);
  }
#undef DDG_CLK
}

// ---------------------------------------------------------------------------------------------------------
// Weight packing: fp32 weights (arbitrary strides) -> per-(n-tile, stage) shared-memory images, bf16 hi [+ lo].
// Blob layout per stage: plane hi then plane lo; plane = [KB/8 chunks][NT rows][8 k] bf16.
// ---------------------------------------------------------------------------------------------------------
template <int PREC>
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin_real,
                                    int Cin_pad, int ntaps, long s_co, long s_ci, long s_tap, int flip_taps, int NT, int KB,
                                    int n_tiles, int stage_offset, int total_stages, long w_batch_stride, long out_batch_stride) {
  // one thread per 16-byte chunk: (ntile, kb, tap, chunk, row); blockIdx.y = batch (per-image operands of the attention GEMMs)
  w += (long)blockIdx.y * w_batch_stride;
  out += (long)blockIdx.y * out_batch_stride;
  const int KCH = KB / 8;
  const long per_stage = (long)KCH * NT;
  const int nkb = Cin_pad / KB;
  const long total = (long)n_tiles * nkb * ntaps * per_stage;
  const long plane_elems = (long)KB * NT;
  const int npl = (PREC == 3) ? 2 : 1;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int row = r % NT; r /= NT;
    const int ch = r % KCH; r /= KCH;
    const int tap = r % ntaps; r /= ntaps;
    const int kb = r % nkb; r /= nkb;
    const int nt = (int)r;
    const int co = nt * NT + row;
    const int tsrc = flip_taps ? (ntaps - 1 - tap) : tap;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = 0.f, b = 0.f;
      const int ci = kb * KB + ch * 8 + 2 * j;
      if (co < Cout) {
        if (ci < Cin_real) a = w[co * s_co + ci * s_ci + tsrc * s_tap];
        if (ci + 1 < Cin_real) b = w[co * s_co + (ci + 1) * s_ci + tsrc * s_tap];
      }
      split_bf16x2(a, b, hi[j], lo[j]);
    }
    const long stage = stage_offset + (long)kb * ntaps + tap;
    __nv_bfloat16* blob = out + ((long)nt * total_stages + stage) * plane_elems * npl;
    uint4* dh = reinterpret_cast<uint4*>(blob + ((long)ch * NT + row) * 8);
    *dh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (PREC == 3) {
      uint4* dl = reinterpret_cast<uint4*>(blob + plane_elems + ((long)ch * NT + row) * 8);
      *dl = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
struct Variant { int msub, nt, kb; };

static int g_small_nt64 = 1;
static thread_local int g_last_info[6] = {0, 0, 0, 0, 0, 1};   // msub, nt, persistent, CTAs, TMA-fed K segments, split-K factor of the last launch
static int g_splitk = getenv("DDG_CONV_NO_SPLITK") ? 0 : 1;
static int g_persist = getenv("DDG_CONV_NO_PERSIST") ? 0 : 1;   // persistent variant (epilogue overlapped with the next tile's mainloop) when tiles > SMs
static int num_sms() {
  static int n = 0;
  if (n == 0) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148; }
  return n;
}
static int g_nt256 = 1;  // N = 256 tiles for Cout % 256 == 0 (96 B/clk of operand reads per MMA instead of 128 B/clk at N = 128)

// Output-channel tile.  N = 256 only when the grid still covers the machine (one wave of 128-row tiles at least).
static int pick_nt(int cout, long m_rows) {
  if (cout <= 16) return 16;
  if (cout <= 64) return 64;
  // (>= 0.8 of a wave: below N = 256 every MMA re-reads a 4 KB A tile for 4 KB of B, right at the 128 B/clk shared-memory limit)
  if (g_nt256 && cout % 256 == 0 && ((m_rows + 127) / 128) * (cout / 256) >= 120) return 256;
  // tiny spatial levels (4x4, 8x8): narrow tiles multiply the CTA count and shorten each CTA's serial MMA chain
  if (g_small_nt64 && cout % 64 == 0 && ((m_rows + 127) / 128) * (cout / 64) <= 148) return 64;
  return 128;
}
static bool valid_nt(int nt) { return nt == 16 || nt == 64 || nt == 128 || nt == 256; }

static int g_nsa_max = getenv("DDG_CONV_NSA") ? atoi(getenv("DDG_CONV_NSA")) : 4;

template <int MSUB, int NT, int KB, int PREC, bool PERSIST = false, bool PROF = false, bool SPLITK = false>
static int launch_conv(ConvDev& d, int n_tiles, cudaStream_t stream) {
  using Cfg = ConvCfg<MSUB, NT, KB, PREC>;
  const size_t a_stage = (size_t)Cfg::NPL * Cfg::KCH * d.win_pitch;
  const size_t fixed = 256 + (size_t)Cfg::NSB * Cfg::B_STAGE + kEpiSlotBytes;
  if (fixed + 2 * a_stage > 227 * 1024) { ddg_set_last_error("conv_tc: shared memory budget exceeded"); return DDG_ERR_UNSUPPORTED; }
  // A ring: as deep as the budget allows, up to 4 stages (1-tap K segments consume a stage in ~12 MMAs: two stages cannot hide the
  // load -> convert -> store latency of the producers)
  int nsa = (int)((227 * 1024 - fixed) / a_stage);
  nsa = nsa > g_nsa_max ? g_nsa_max : nsa;
  if (nsa < 2) nsa = 2;
  if (nsa > 4) nsa = 4;
  d.nsa = nsa;
  const size_t smem = fixed + (size_t)nsa * a_stage;
  auto kern = conv_tc_kernel<MSUB, NT, KB, PREC, PERSIST, PROF, SPLITK>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  dim3 grid((d.Mtotal + Cfg::MT - 1) / Cfg::MT, n_tiles);
  if (d.tile2d) grid = dim3(d.N * d.tiles_x * d.tiles_y, n_tiles);
  if (d.batch_rows > 0) grid = dim3((d.batch_rows + Cfg::MT - 1) / Cfg::MT, n_tiles, d.Mtotal / d.batch_rows);
  d.tiles_m = (int)grid.x;
  d.n_tiles = n_tiles;
  g_last_info[0] = MSUB; g_last_info[1] = NT; g_last_info[2] = PERSIST ? 1 : 0;
  g_last_info[3] = (int)(grid.x * grid.y * grid.z);
  if (PERSIST) {
    const long total = (long)grid.x * n_tiles;
    grid = dim3((unsigned)(total < num_sms() ? total : num_sms()));
    g_last_info[3] = (int)grid.x;
    d.ksplit = 1;
  } else if (SPLITK) {
    grid.z = d.ksplit;                    // rank = blockIdx.z: every writer rank is dispatched before the reducing rank (z = ksplit - 1)
    g_last_info[3] *= d.ksplit;
  }
  launch_pdl(kern, grid, dim3(PERSIST ? kThreadsPersist : kThreads), smem, stream, d);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

// fp32 PNHWC -> pre-split bf16 planes [plane][N][C/8][H+2][W+2][8] (interior only).  One thread per (pixel, 8-channel chunk), pixels
// fastest: 32-byte reads (one sector each), 16-byte writes contiguous across the warp.
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, uint8_t* __restrict__ planes, int N, int H, int W, int C,
                                                          int nplanes, long plane_bytes) {
  const int C8 = C / 8;
  const long total = (long)N * C8 * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int ch = (int)(r % C8);
    const int n = (int)(r / C8);
    const size_t pix = (size_t)(n * (H + 2) + h + 1) * (W + 2) + (w + 1);
    const float4* src = reinterpret_cast<const float4*>(x + pix * C + ch * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    uint4 hi, lo;
    split_bf16x2(a.x, a.y, hi.x, lo.x); split_bf16x2(a.z, a.w, hi.y, lo.y);
    split_bf16x2(b.x, b.y, hi.z, lo.z); split_bf16x2(b.z, b.w, hi.w, lo.w);
    const size_t off = (((size_t)n * C8 + ch) * (size_t)(H + 2) * (W + 2) + (size_t)(h + 1) * (W + 2) + (w + 1)) * 16;
    *reinterpret_cast<uint4*>(planes + off) = hi;
    if (nplanes == 2) *reinterpret_cast<uint4*>(planes + plane_bytes + off) = lo;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}
static int g_use_tma = getenv("DDG_CONV_NO_TMA") ? 0 : 1;

// planes [N][Ct/8][Hp][Wp][8] bf16 as a 4-D tensor (innermost first: Wp*8, Hp, Ct/8, N); box = 80 x rows x 4 chunks x 1
static bool encode_plane_map(CUtensorMap* m, const void* base, int N, int Hp, int Wp, int Ct, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  // [n][chunk][y][x * 8 + c]: the 8 channels of a chunk and the image columns are one contiguous dimension, so a halo window row is
  // one 160-byte box row (the first version had a separate 8-element inner dimension: 16-byte rows, ten times the TMA requests, and
  // was slower than converting fp32 in the producer warps)
  const cuuint64_t dims[4] = {(cuuint64_t)Wp * 8, (cuuint64_t)Hp, (cuuint64_t)(Ct / 8), (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)Wp * 16, (cuuint64_t)Hp * Wp * 16, (cuuint64_t)(Ct / 8) * Hp * Wp * 16};
  const cuuint32_t box[4] = {80, (cuuint32_t)box_rows, 4, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ddg

using namespace ddg;

extern "C" long ddg_planes_bytes(int N, int H, int W, int C) { return (long)N * (C / 8) * (H + 2) * (W + 2) * 16; }

extern "C" int ddg_split_planes(const float* x, void* planes, int N, int H, int W, int C, int nplanes, cudaStream_t stream) {
  if (!x || !planes || C % 8 != 0 || nplanes < 1 || nplanes > 2) { ddg_set_last_error("split_planes: bad args"); return DDG_ERR_ARG; }
  const long total = (long)N * (C / 8) * H * W;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  split_planes_kernel<<<(int)blocks, 256, 0, stream>>>(x, (uint8_t*)planes, N, H, W, C, nplanes, ddg_planes_bytes(N, H, W, C));
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_conv_tile_n(int cout, long m_rows) { return pick_nt(cout, m_rows); }
extern "C" int ddg_conv_last_launch_tma(void) { return g_last_info[4]; }
extern "C" int ddg_conv_last_launch_ksplit(void) { return g_last_info[5]; }
extern "C" int ddg_conv_last_launch_info(int* msub, int* nt, int* persistent, int* grid_ctas) {
  if (msub) *msub = g_last_info[0];
  if (nt) *nt = g_last_info[1];
  if (persistent) *persistent = g_last_info[2];
  if (grid_ctas) *grid_ctas = g_last_info[3];
  return DDG_OK;
}
// tuning switches (bring-up / A-B measurements): bit 0 = N=256 tiles, bit 1 = no N=64 on tiny levels, bit 2 = no persistent variant
extern "C" int ddg_conv_set_nt256(int on) {
  const int old = g_nt256;
  g_nt256 = on & 1; g_small_nt64 = (on >> 1) & 1 ? 0 : 1; g_persist = (on >> 2) & 1 ? 0 : 1;
  return old;
}

extern "C" long ddg_conv_packed_bytes(int cout, int total_stages, int kb, int precision, int nt) {
  if (!valid_nt(nt)) return -1;
  const int n_tiles = (cout + nt - 1) / nt;
  return (long)n_tiles * total_stages * kb * nt * 2 * (precision == 3 ? 2 : 1);
}

extern "C" int ddg_conv_pack_weights(const float* w, void* out, int cout, int cin_real, int cin_pad, int ntaps, long s_co,
                                     long s_ci, long s_tap, int flip_taps, int kb, int stage_offset, int total_stages,
                                     int precision, int nt, int batch, long w_batch_stride, cudaStream_t stream) {
  if (!w || !out || cin_pad % kb != 0 || (kb != 32 && kb != 64) || !valid_nt(nt)) { ddg_set_last_error("pack_weights: bad args"); return DDG_ERR_ARG; }
  const int n_tiles = (cout + nt - 1) / nt;
  const long total = (long)n_tiles * (cin_pad / kb) * ntaps * (kb / 8) * nt;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < 148 * 16 ? (total + threads - 1) / threads : 148 * 16);
  if (batch < 1) batch = 1;
  const long out_bs = ddg_conv_packed_bytes(cout, total_stages, kb, precision, nt) / 2;  // in bf16 elements
  dim3 grid(blocks, batch);
  if (precision == 3)
    pack_weights_kernel<3><<<grid, threads, 0, stream>>>(w, (__nv_bfloat16*)out, cout, cin_real, cin_pad, ntaps, s_co, s_ci,
                                                        s_tap, flip_taps, nt, kb, n_tiles, stage_offset, total_stages, w_batch_stride, out_bs);
  else
    pack_weights_kernel<1><<<grid, threads, 0, stream>>>(w, (__nv_bfloat16*)out, cout, cin_real, cin_pad, ntaps, s_co, s_ci,
                                                        s_tap, flip_taps, nt, kb, n_tiles, stage_offset, total_stages, w_batch_stride, out_bs);
  DDG_CHECK_LAUNCH();
  return DDG_OK;
}

extern "C" int ddg_conv2d_fwd(const ddg_conv_desc* c, cudaStream_t stream) {
  if (!c || c->nsrc < 1 || c->nsrc > DDG_CONV_MAX_SRC || !c->out || !c->wpack) { ddg_set_last_error("conv2d_fwd: bad args"); return DDG_ERR_ARG; }
  const int KB = c->kb;
  if (KB != 32) { ddg_set_last_error("conv2d_fwd: kb must be 32"); return DDG_ERR_UNSUPPORTED; }
  ConvDev d{};
  d.nsrc = c->nsrc;
  bool window = false;
  int total_stages = 0;
  for (int s = 0; s < c->nsrc; ++s) {
    const ddg_conv_src& S = c->src[s];
    if (!S.x || S.C % KB != 0 || S.ntaps < 1 || S.ntaps > 9) { ddg_set_last_error("conv2d_fwd: bad source"); return DDG_ERR_ARG; }
    if ((S.scale == nullptr) != (S.shift == nullptr)) { ddg_set_last_error("conv2d_fwd: scale/shift must come together"); return DDG_ERR_ARG; }
    if (S.ntaps > 1) window = true;
    total_stages += (S.C / KB) * S.ntaps;
  }
  d.window = window ? 1 : 0;
  d.N = c->N; d.Hout = c->Hout; d.Wout = c->Wout;
  d.Hp = c->Hp; d.Wp = c->Wp;
  if (window) {
    if (d.Hp < d.Hout + 2 || d.Wp < d.Wout + 2) { ddg_set_last_error("conv2d_fwd: padded dims too small"); return DDG_ERR_ARG; }
    d.Mtotal = d.N * d.Hp * d.Wp;
    d.margin = d.Wp + 1;
  } else {
    d.Mtotal = d.N * d.Hout * d.Wout;
    d.margin = 0;
  }
  for (int s = 0; s < c->nsrc; ++s) {
    const ddg_conv_src& S = c->src[s];
    SrcDev& D = d.src[s];
    D.x = S.x; D.scale = S.scale; D.shift = S.shift; D.C = S.C; D.act = S.act; D.ntaps = S.ntaps; D.padded = S.padded;
    D.pitch = S.pitch > 0 ? S.pitch : S.C;
    D.ss_stride = S.ss_stride > 0 ? S.ss_stride : S.C;
    if (D.pitch % 4 != 0 || D.ss_stride % 4 != 0 || (((uintptr_t)S.x) & 15) != 0) { ddg_set_last_error("conv2d_fwd: sources must be 16-byte aligned with pitch % 4 == 0"); return DDG_ERR_ARG; }
    for (int t = 0; t < S.ntaps; ++t) {
      if (window) {
        if (!S.padded) { ddg_set_last_error("conv2d_fwd: window mode needs padded sources"); return DDG_ERR_ARG; }
        D.tapoff[t] = d.margin + S.tap_dr[t] * d.Wp + S.tap_ds[t];
      } else {
        D.tapoff[t] = 0;
      }
    }
  }
  d.wpack = (const __nv_bfloat16*)c->wpack;
  d.Cout = c->Cout;
  d.total_stages = total_stages;
  d.bias = c->bias; d.addvec = c->addvec; d.addvec_stride = c->addvec_stride;
  d.res = c->res; d.out_scale = c->out_scale; d.out_act = c->out_act;
  d.out = c->out; d.out_mode = c->out_mode; d.out_C = c->out_C > 0 ? c->out_C : c->Cout;
  d.stats = c->stats;
  d.batch_rows = 0;
  d.w_batch_stride = 0;
  d.prof = (long long*)c->debug_prof;
  if (c->batch_rows > 0) {
    if (window || d.Mtotal % c->batch_rows != 0) { ddg_set_last_error("conv2d_fwd: batched mode needs a 1x1 problem with Mtotal % batch_rows == 0"); return DDG_ERR_ARG; }
    d.batch_rows = c->batch_rows;
    d.w_batch_stride = ddg_conv_packed_bytes(c->Cout, total_stages, KB, c->precision == 1 ? 1 : 3, c->nt);
  }
  if (d.out_mode != 2 && (d.out_C % 4 != 0)) { ddg_set_last_error("conv2d_fwd: NHWC output pitch must be a multiple of 4"); return DDG_ERR_ARG; }

  const int nt = c->nt;
  if (!valid_nt(nt)) { ddg_set_last_error("conv2d_fwd: desc.nt must be the tile width the weights were packed with (16/64/128/256)"); return DDG_ERR_ARG; }
  const int n_tiles = (c->Cout + nt - 1) / nt;
  if (d.out_mode != 2 && (c->Cout % (nt >= 64 ? 32 : 16) != 0)) { ddg_set_last_error("conv2d_fwd: Cout must be a multiple of the epilogue chunk for NHWC output"); return DDG_ERR_UNSUPPORTED; }
  // 2-D tiles (16*msub rows x 8 columns) whenever the geometry allows: no padded-space waste, window independent of W
  bool tile2d = window && !c->force_linear && d.Hp == d.Hout + 2 && d.Wp == d.Wout + 2 && d.Wout % 8 == 0 && d.Hout % 16 == 0 &&
                c->batch_rows == 0;
  for (int s = 0; s < c->nsrc && tile2d; ++s)
    for (int t = 0; t < c->src[s].ntaps; ++t)
      if (c->src[s].tap_dr[t] < -1 || c->src[s].tap_dr[t] > 1 || c->src[s].tap_ds[t] < -1 || c->src[s].tap_ds[t] > 1) tile2d = false;
  // M sub-tiles: two accumulators per CTA when the grid still fills the machine
  int msub = c->msub;
  if (msub == 0) {
    const long tiles2 = tile2d ? (long)d.N * (d.Hout / 32) * (d.Wout / 8) * n_tiles : ((long)d.Mtotal + 255) / 256 * n_tiles;
    msub = (tiles2 >= 222) ? 2 : 1;   // >= 1.5 waves of 256-row tiles
    if (d.batch_rows > 0 && d.batch_rows % 256 != 0) msub = 1;
  }
  if (tile2d && d.Hout % (16 * msub) != 0) msub = 1;
  // Persistent variant: more tiles than SMs, so every CTA runs >= 2 tiles back to back and the epilogue of tile i overlaps
  // the mainloop of tile i+1 (two accumulator sets in TMEM: msub * nt <= 256 columns each).
  bool persist = false;
  if (g_persist && c->batch_rows == 0) {
    auto tiles_for = [&](int ms) -> long {
      return (tile2d ? (long)d.N * (d.Hout / (16 * ms)) * (d.Wout / 8) : ((long)d.Mtotal + 128 * ms - 1) / (128 * ms)) * n_tiles;
    };
    int ms = msub;
    if (ms * nt > 256) ms = 1;
    const bool have = (ms == 2 && (nt == 128 || nt == 64 || nt == 16)) || (ms == 1 && (nt == 256 || nt == 128));
    if (have && tiles_for(ms) > num_sms()) { persist = true; msub = ms; }
  }
  const int MT = 128 * msub;
  d.tile2d = tile2d ? 1 : 0;
  d.a_sbo = 128;
  d.sub_stride = 128 * 16;
  if (tile2d) {
    d.tiles_x = d.Wout / 8;
    d.tiles_y = d.Hout / (16 * msub);
    d.a_sbo = 10 * 16;
    d.sub_stride = 16 * 10 * 16;
    d.win_rows = (16 * msub + 2) * 10;
    for (int s = 0; s < c->nsrc; ++s)
      for (int t = 0; t < c->src[s].ntaps; ++t) d.src[s].tapoff[t] = (c->src[s].tap_dr[t] + 1) * 10 + (c->src[s].tap_ds[t] + 1);
  } else {
    d.win_rows = window ? MT + 2 * d.margin : MT;
  }
  int rows = d.win_rows;
  rows += (10 - (rows & 7)) & 7;                    // pitch = 32 (mod 128) bytes: the 4 chunks x 2 rows of a 16-byte store phase hit 8 distinct bank groups
  d.win_pitch = rows * 16;
  const int prec = c->precision == 1 ? 1 : 3;
  // TMA-fed K segments: prologue-free sources that come with pre-split planes, 2-D tiling only (the halo window is a 4-D box)
  d.tma_pitch = d.win_rows * 16;
  for (int s = 0; s < c->nsrc; ++s) {
    const ddg_conv_src& S = c->src[s];
    d.src[s].tma = 0;
    d.src[s].chunk0 = 0;
    if (!g_use_tma || !tile2d || !S.planes || S.scale || S.act != 0 || S.planes_C % 8 != 0 || S.planes_c0 % 8 != 0) continue;
    const long pb = ddg_planes_bytes(d.N, d.Hout, d.Wout, S.planes_C);
    bool ok = encode_plane_map(&d.tmap[s][0], S.planes, d.N, d.Hout + 2, d.Wout + 2, S.planes_C, 16 * msub + 2);
    if (ok && prec == 3) ok = encode_plane_map(&d.tmap[s][1], (const uint8_t*)S.planes + pb, d.N, d.Hout + 2, d.Wout + 2, S.planes_C, 16 * msub + 2);
    if (ok) { d.src[s].tma = 1; d.src[s].chunk0 = S.planes_c0 / 8; }
  }
  g_last_info[4] = 0;
  for (int s = 0; s < c->nsrc; ++s) g_last_info[4] += d.src[s].tma;
  if (c->out_mode == 0 && c->zero_border) {   // fresh output buffer: clear its one-pixel frame first (same stream)
    const int rc = ddg_zero_border(c->out, c->N, c->Hout, c->Wout, d.out_C, stream);
    if (rc != DDG_OK) return rc;
  }
  if ((long)c->N * (c->Hout + 2) * (c->Wout + 2) >= (1L << 31)) { ddg_set_last_error("conv2d_fwd: more than 2^31 output pixels"); return DDG_ERR_UNSUPPORTED; }
  d.out_planes = (c->out_mode == 0) ? c->out_planes : nullptr;
  d.out_plane_bytes = ddg_planes_bytes(d.N, d.Hout, d.Wout, d.out_C);

  // split-K for grids that leave most of the machine idle (4x4 / 8x8 levels: a few dozen one-tile CTAs, each a serial chain of
  // K/16 * 3 MMAs): ksplit CTAs per tile, all resident at once (tiles * ksplit <= SMs), reduction through the caller's workspace.
  d.ksplit = 1;
  if (g_splitk && !persist && msub == 1 && (nt == 64 || nt == 128) && c->splitk_ws != nullptr && c->batch_rows == 0 && c->debug_prof == nullptr) {
    const long tiles = (tile2d ? (long)d.N * (d.Hout / (16 * msub)) * (d.Wout / 8) : ((long)d.Mtotal + MT - 1) / MT) * n_tiles;
    int nkb = 0;
    for (int s = 0; s < c->nsrc; ++s) nkb += c->src[s].C / KB;
    for (int cand = 4; cand >= 2 && d.ksplit == 1; cand -= 2) {
      if (tiles * cand > num_sms() || nkb < 2 * cand) continue;
      const long need = 4096 + tiles * (cand - 1) * (long)MT * nt * 4;
      if (tiles * 4 > 4096 || need > c->splitk_ws_bytes) continue;
      // balance by B stages (a K block of a 3x3 segment carries nine of them, one of a 1x1 segment)
      int split[5] = {0, 0, 0, 0, 0}, r = 1, kbi = 0;
      long acc_st = 0;
      for (int s = 0; s < c->nsrc && r < cand; ++s)
        for (int kb = 0; kb < c->src[s].C / KB && r < cand; ++kb, ++kbi) {
          acc_st += c->src[s].ntaps;
          if (acc_st * cand >= (long)total_stages * r) split[r++] = kbi + 1;
        }
      split[cand] = nkb;
      bool ok = (r == cand);
      for (int q = 0; q < cand && ok; ++q) ok = split[q + 1] > split[q];
      if (!ok) continue;
      d.ksplit = cand;
      for (int q = 0; q <= cand; ++q) d.kb_split[q] = split[q];
      d.ws_flag = reinterpret_cast<unsigned*>(c->splitk_ws);
      d.ws_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(c->splitk_ws) + 4096);
    }
  }
  g_last_info[5] = d.ksplit;

  // tuning builds with cycle counters (tools/conv_prof.py): the three shapes that dominate a generator forward
  if (d.prof != nullptr && prec == 3) {
    if (persist && nt == 128 && msub == 2) return launch_conv<2, 128, 32, 3, true, true>(d, n_tiles, stream);
    if (persist && nt == 256) return launch_conv<1, 256, 32, 3, true, true>(d, n_tiles, stream);
    if (!persist && nt == 256 && msub == 1) return launch_conv<1, 256, 32, 3, false, true>(d, n_tiles, stream);
    if (!persist && nt == 64 && msub == 1) return launch_conv<1, 64, 32, 3, false, true>(d, n_tiles, stream);
    if (!persist && nt == 128 && msub == 1) return launch_conv<1, 128, 32, 3, false, true>(d, n_tiles, stream);
    d.prof = nullptr;   // no counter build of this variant: run the production kernel
  } else {
    d.prof = nullptr;
  }
#define DDG_LAUNCH_P(MS, NTV, PR) return launch_conv<MS, NTV, 32, PR, true>(d, n_tiles, stream)
  if (persist) {
    if (prec == 3) {
      if (nt == 256) DDG_LAUNCH_P(1, 256, 3);
      if (nt == 128) { if (msub == 2) DDG_LAUNCH_P(2, 128, 3); else DDG_LAUNCH_P(1, 128, 3); }
      if (nt == 64) DDG_LAUNCH_P(2, 64, 3);
      if (nt == 16) DDG_LAUNCH_P(2, 16, 3);
    } else {
      if (nt == 256) DDG_LAUNCH_P(1, 256, 1);
      if (nt == 128) { if (msub == 2) DDG_LAUNCH_P(2, 128, 1); else DDG_LAUNCH_P(1, 128, 1); }
      if (nt == 64) DDG_LAUNCH_P(2, 64, 1);
      if (nt == 16) DDG_LAUNCH_P(2, 16, 1);
    }
  }
#undef DDG_LAUNCH_P
  if (d.ksplit > 1) {       // (msub == 1, nt 64 / 128: see the eligibility test above)
    if (prec == 3) { if (nt == 64) return launch_conv<1, 64, 32, 3, false, false, true>(d, n_tiles, stream); return launch_conv<1, 128, 32, 3, false, false, true>(d, n_tiles, stream); }
    if (nt == 64) return launch_conv<1, 64, 32, 1, false, false, true>(d, n_tiles, stream);
    return launch_conv<1, 128, 32, 1, false, false, true>(d, n_tiles, stream);
  }
#define DDG_LAUNCH(MS, NTV, PR) return launch_conv<MS, NTV, 32, PR>(d, n_tiles, stream)
  if (prec == 3) {
    if (nt == 256) { if (msub == 2) DDG_LAUNCH(2, 256, 3); else DDG_LAUNCH(1, 256, 3); }
    if (nt == 128) { if (msub == 2) DDG_LAUNCH(2, 128, 3); else DDG_LAUNCH(1, 128, 3); }
    if (nt == 64) { if (msub == 2) DDG_LAUNCH(2, 64, 3); else DDG_LAUNCH(1, 64, 3); }
    if (nt == 16) { if (msub == 2) DDG_LAUNCH(2, 16, 3); else DDG_LAUNCH(1, 16, 3); }
  } else {
    if (nt == 256) { if (msub == 2) DDG_LAUNCH(2, 256, 1); else DDG_LAUNCH(1, 256, 1); }
    if (nt == 128) { if (msub == 2) DDG_LAUNCH(2, 128, 1); else DDG_LAUNCH(1, 128, 1); }
    if (nt == 64) { if (msub == 2) DDG_LAUNCH(2, 64, 1); else DDG_LAUNCH(1, 64, 1); }
    if (nt == 16) { if (msub == 2) DDG_LAUNCH(2, 16, 1); else DDG_LAUNCH(1, 16, 1); }
  }
#undef DDG_LAUNCH
  ddg_set_last_error("conv2d_fwd: no kernel variant");
  return DDG_ERR_UNSUPPORTED;
}
