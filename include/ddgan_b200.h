/* ddgan_b200 -- C ABI of the B200-native DDGAN hot path (libddgan_b200.so).
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t; none allocates, synchronises or throws.
 * Return value: 0 = ok, <0 = error (ddg_last_error() gives the text).  All tensors are fp32 unless stated.
 * The reference interface each function replaces is cited as file:line under /root/reference.
 *
 * Internal activation layout ("PNHWC"): [N][H+2][W+2][C] fp32 with an all-zero one-pixel border, C a multiple of 32.
 * Public operator surface (score_sde.op.*) is NCHW like the reference.
 */
#ifndef DDGAN_B200_H
#define DDGAN_B200_H

#include <stdint.h>
#ifndef __CUDACC__
typedef struct CUstream_st* cudaStream_t;
#else
#include <cuda_runtime.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

const char* ddg_last_error(void);
int ddg_version(void);

/* ---- score_sde/op/upfirdn2d.cpp:20-31 (upfirdn2d_op.upfirdn2d) + upfirdn2d_kernel.cu:211-371 -------------------
 * x: [planes][in_h][in_w] (planes = N*C, minor = 1 as the Python layer always passes, upfirdn2d.py:107),
 * k: [kh][kw] on device.  out: [planes][out_h][out_w], out_h = (in_h*up_y + pad_y0 + pad_y1 - kh)/down_y + 1.
 * The adjoint (UpFirDn2dBackward, upfirdn2d.py:27-68) and the double-backward (:70-93) are the same entry with
 * up/down swapped and the g_pad arithmetic of upfirdn2d.py:119-122 done by the caller. */
int ddg_upfirdn2d(const float* x, const float* k, float* out, long planes, int in_h, int in_w, int kh, int kw,
                  int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                  cudaStream_t stream);
int ddg_upfirdn2d_out_size(int in_size, int up, int down, int pad0, int pad1, int ksize);

/* ---- score_sde/op/fused_bias_act.cpp:18-28 (fused.fused_bias_act), fused_bias_act_kernel.cu:20-101 --------------
 * y = f(x + b[(i / step_b) % size_b]) * scale; act=3 (leaky relu): grad=0 -> f = lrelu(alpha); grad=1 -> gate by
 * ref > 0; grad=2 -> 0.  act=1 linear.  b / ref may be NULL ("empty tensor" in the reference). */
int ddg_fused_bias_act(const float* x, const float* b, const float* ref, float* y, long n, int step_b, int size_b,
                       int act, int grad, float alpha, float scale, cudaStream_t stream);
/* 16-bit I/O variants of the two operators (the reference dispatches over AT_DISPATCH_FLOATING_TYPES_AND_HALF,
 * upfirdn2d_kernel.cu:313 / fused_bias_act_kernel.cu:79): dtype 1 = fp16, 2 = bf16 for x / ref / out; taps and bias stay fp32,
 * accumulation is fp32.  up / down / pad apply to both axes as in the Python surface (upfirdn2d.py:153-164). */
int ddg_upfirdn2d_lp(const void* x, const float* k, void* out, long planes, int in_h, int in_w, int kh, int kw, int up, int down,
                     int pad0, int pad1, int dtype, cudaStream_t stream);
int ddg_fused_bias_act_lp(const void* x, const float* b, const void* ref, void* y, long n, int step_b, int size_b, int act, int grad,
                          float alpha, float scale, int dtype, cudaStream_t stream);
/* grad_bias of FusedLeakyReLUFunctionBackward (fused_act.py:42-47): out[c] = sum over n, spatial of g[n][c][...] */
int ddg_channel_sum(const float* g, float* out, int N, int C, int inner, cudaStream_t stream);

/* ---- nn.GroupNorm / AdaptiveGroupNorm (layerspp.py:46-63, :100; ncsnpp_generator_adagn.py:264), NCHW ------------
 * y = act(gamma[n,c] * (x - mean)/sqrt(var+eps) + beta[n,c]); gamma/beta: per-(n,c) (stride C) when per_sample=1,
 * per-channel when 0, NULL = identity.  mean/rstd [N*G] are written (needed by the backward). */
int ddg_groupnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int N,
                      int C, int HW, int G, float eps, int per_sample, int act, cudaStream_t stream);
/* backward: dx, and (optional) dgamma/dbeta reduced over HW per (n,c) -- the caller reduces over n for affine GN. */
int ddg_groupnorm_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, float* dx, float* dgamma_nc, float* dbeta_nc, int N, int C, int HW, int G,
                      int per_sample, int act, cudaStream_t stream);

/* ---- fused small kernels -------------------------------------------------------------------------------------
 * layers.py:475-486 get_timestep_embedding: out[n][j] = sin|cos(t[n] * exp(-j ln(max_pos)/(half-1))) */
int ddg_timestep_embedding(const int64_t* t, float* out, int N, int dim, float max_positions, cudaStream_t stream);
/* y[n][j] = act_out( sum_k act_in(x[n][k]) * W[j][k] + b[j] ), nn.Linear layout W [out][in].  pixel_norm=1 applies
 * ncsnpp_generator_adagn.py:51-56 (x / sqrt(mean(x^2)+1e-8)) to the input row first.  Used for the z-mapping MLP
 * (:271-277), the temb MLP (:301-303), AdaGN style projections (layerspp.py:57) and Dense_0 (:298-299). */
int ddg_linear(const float* x, const float* W, const float* b, float* y, int N, int K, int J, int ldx, int ldy, int act_in,
               int act_out, int pixel_norm, cudaStream_t stream);
/* A whole small MLP on [N, K0] rows as ONE kernel: y = L_{n-1}(act(... act(L_0(norm(x))) ...)), L_i(v) = W_i v + b_i with W_i
 * [dims[i+1]][dims[i]] (nn.Linear layout).  pixel_norm applies ncsnpp_generator_adagn.py:51-56 to the input row; `act` is
 * applied between layers (not after the last one: consumers fold it into their prologue).  Replaces the z-mapping network
 * (PixelNorm + 1 + n_mlp dense layers, ncsnpp_generator_adagn.py:271-277) and the time-embedding MLP (:301-303).
 * Limits: nlayers <= DDG_MLP_MAX_LAYERS, every dims[i] <= 1024. */
#define DDG_MLP_MAX_LAYERS 8
typedef struct {
  const float* W[DDG_MLP_MAX_LAYERS];
  const float* b[DDG_MLP_MAX_LAYERS];
  int dims[DDG_MLP_MAX_LAYERS + 1];
  int nlayers;
  int pixel_norm;
  int act;
} ddg_mlp_desc;
int ddg_mlp_rows(const float* x, int ldx, float* y, int ldy, int N, const ddg_mlp_desc* desc, cudaStream_t stream);
/* ddgan.py:110-126 q_sample_pairs with injected noise: x_t = a_cum[t] x0 + s_cum[t] n0 ; x_tp1 = a[t+1] x_t + s[t+1] n1 */
int ddg_q_sample_pairs(const float* x0, const float* noise_xt, const float* noise_xtp1, const int64_t* t, const float* a_s_cum,
                       const float* sigmas_cum, const float* a_s, const float* sigmas, float* x_t, float* x_tp1, int N,
                       long per_sample, cudaStream_t stream);
/* ddgan.py:152-169 / test_ddgan.py:96-113 sample_posterior with injected noise */
int ddg_sample_posterior(const float* x0, const float* x_t, const float* noise, const int64_t* t, const float* coef1,
                         const float* coef2, const float* logvar, float* out, int N, long per_sample, cudaStream_t stream);

/* ---- internal layout helpers ---------------------------------------------------------------------------------
 * NCHW sources (up to 2, concatenated along C: discriminator.py:138 cat(x, x_t)) -> PNHWC with C padded to Cpad. */
int ddg_nchw_to_pnhwc(const float* a, int Ca, const float* b, int Cb, float* out, int N, int H, int W, int Cpad,
                      float scale, float shift, cudaStream_t stream);
int ddg_pnhwc_to_nchw(const float* x, float* out, int N, int H, int W, int C, int Cpitch, int padded, cudaStream_t stream);
/* AdaGN / GN prologue coefficients from per-(n,c) sums written by the conv epilogue:
 *   scale[n,c] = gamma*rstd ; shift[n,c] = beta - mean*gamma*rstd, groups over the concatenation of up to two
 *   sources (ncsnpp_generator_adagn.py:367 torch.cat([h, hs.pop()])).  gamma/beta as in ddg_groupnorm_fwd. */
int ddg_gn_prepare(const double* stats_a, int Ca, const double* stats_b, int Cb, const float* gamma, const float* beta,
                   int gb_stride, int per_sample, float* scale, float* shift, int N, int HW, int G, float eps,
                   cudaStream_t stream);
/* backward of ddg_gn_prepare (training): given d(scale), d(shift) [N][C] returns d(stats) [N][C][2] (double), and the per-sample
 * d(gamma), d(beta) [N][C] (sum over N on the caller's side for a shared affine).  Same group statistics as the forward. */
int ddg_gn_prepare_bwd(const double* stats, const float* gamma, int gb_stride, int per_sample, const float* dscale,
                       const float* dshift, double* dstats, float* dgamma, float* dbeta, int N, int C, int HW, int G, float eps,
                       cudaStream_t stream);
/* FIR resampling on PNHWC with the AdaGN+SiLU prologue fused on load (layerspp.py:279-293): [1,3,3,1] (x) [1,3,3,1]
 * mode 1: up x2 (upsample_2d, up_or_down_sampling.py:200-228), mode 2: down x2 (downsample_2d, :231-261),
 * mode 3: pad (2,2) FIR (H -> H+1) written space-to-depth for the stride-2 conv of conv_downsample_2d (:149-183):
 *         out [N][Ho+3][Wo+3][4*C -> Cout_pitch], Ho = H/2, cell (i,j) channel (py*2+px)*C + c = fir[2i+py][2j+px][c].
 * mode 4: adjoint of mode 3 (x is the space-to-depth gradient [N][H/2+3][W/2+3][in_pitch], out is [N][H+2][W+2][C]).
 * gain multiplies the taps (the adjoint of up x2 is down x2 with gain 4, the adjoint of down x2 is up x2 with gain 1/4). */
int ddg_fir_pnhwc(const float* x, const float* scale, const float* shift, int act, float* out, int N, int H, int W, int C,
                  int mode, int out_pitch, float gain, cudaStream_t stream);
/* minibatch stddev feature (discriminator.py:150-158) from a PNHWC tensor -> PNHWC [N][H+2][W+2][Cpad] channel 0 */
int ddg_minibatch_stddev(const float* x, float* out, int N, int H, int W, int C, int Cpad, int group, cudaStream_t stream);
/* out[n][c] = sum_{h,w} act(x[n][h][w][c]) over the interior of a PNHWC tensor (discriminator.py:163-165) */
int ddg_spatial_sum(const float* x, float* out, int N, int H, int W, int C, int act, cudaStream_t stream);
/* zero the one-pixel frame of a PNHWC buffer [N][H+2][W+2][C] (C % 4 == 0): kernels write interiors only, so a fresh buffer needs
 * just its border cleared to serve as the zero padding of the next 3x3 conv (ncsnpp convs use padding=1, layerspp.py:33) */
/* Programmatic dependent launch (CUDA PDL) for the kernels of the generator / discriminator forward path: a kernel's CTAs are
 * scheduled and set up while its predecessor on the stream drains, and block (griddepcontrol.wait) before their first global access.
 * Only effective in a library built with -DDDG_ENABLE_PDL (the default build leaves the griddepcontrol instructions out: no gain
 * measured inside the captured loops, and the bare wait instruction cost ~2 %); off by default even then (DDG_PDL=1); returns the
 * previous setting.  No reference counterpart: the reference's
 * launches are PyTorch's. */
int ddg_set_pdl(int on);
int ddg_zero_border(float* buf, int N, int H, int W, int C, cudaStream_t stream);
/* row softmax of the attention logits (layerspp.py:116-118): p[r][0:T] = softmax(s[r][0:T]), p[r][T:ldp] = 0 */
int ddg_softmax_rows(const float* s, float* p, long rows, int T, int lds, int ldp, cudaStream_t stream);
/* its backward (training path of the attention core): ds[r][j] = scale * p[r][j] * (dp[r][j] - sum_k p[r][k] dp[r][k]), ds[r][T:ld] = 0 */
int ddg_softmax_rows_bwd(const float* p, const float* dp, float* ds, long rows, int T, int ld, float scale, cudaStream_t stream);

/* ---- implicit-GEMM convolution on tcgen05 / TMEM (replaces nn.Conv2d -> cuDNN: layers.py:114-138,
 *      dense_layer.py:73-80, NIN layers.py:489-512, up_or_down_sampling.py:56,183) -------------------------------- */
#define DDG_CONV_MAX_SRC 3
typedef struct {
  const float* x;      /* PNHWC (padded=1) or NHWC (padded=0) source, channel pitch C */
  const float* scale;  /* [N][C] prologue scale or NULL */
  const float* shift;  /* [N][C] prologue shift or NULL */
  int C;               /* channels contributed, multiple of kb */
  int pitch;           /* row pitch of x in floats (0 = C): lets a source be a channel slice of a wider tensor */
  int ss_stride;       /* row pitch of scale/shift in floats (0 = C) */
  int act;             /* prologue activation: 0 none, 1 SiLU, 2 LeakyReLU(0.2) */
  int ntaps;           /* taps this source contributes (9 = 3x3, 4 = 2x2, 1 = 1x1) */
  int padded;
  int8_t tap_dr[9];    /* row / column offset of every tap relative to the output position */
  int8_t tap_ds[9];
  /* Optional pre-split operand planes of the SAME tensor (sources without a prologue only): bf16 [planes][N][Ct/8][H+2][W+2][8],
   * plane 0 = rn_bf16(x), plane 1 (precision 3) = rn_bf16(x - plane 0), zero border, written by a conv epilogue
   * (ddg_conv_desc.out_planes) or by ddg_split_planes.  When given and the conv runs the 2-D tiling, the A operand of this K
   * segment is fetched by the TMA engine (cp.async.bulk.tensor 5-D boxes) straight into the UMMA layout -- no conversion work
   * in the kernel; otherwise the fp32 path is used.  planes_C = Ct (channels of the planes tensor), planes_c0 = first channel of
   * this segment inside it (multiple of 8). */
  const void* planes;
  int planes_C, planes_c0;
} ddg_conv_src;

typedef struct {
  ddg_conv_src src[DDG_CONV_MAX_SRC]; /* K segments, consumed in order (concat inputs / fused 1x1 skip conv) */
  int nsrc;
  const void* wpack;   /* ddg_conv_pack_weights output */
  int kb;              /* K block (32) */
  int nt;              /* output-channel tile the weights were packed with (ddg_conv_tile_n) */
  int N, Hout, Wout;   /* output image size */
  int Hp, Wp;          /* padded input space (any source with ntaps > 1): Hout+2 x Wout+2 for 3x3 pad 1 */
  int Cout;
  const float* bias;   /* [Cout] or NULL */
  const float* addvec; /* [N][addvec_stride] per-sample per-channel add (Dense_0(temb)) or NULL */
  int addvec_stride;
  const float* res;    /* residual, same layout as out, or NULL */
  float out_scale;     /* applied after the residual add (1/sqrt2 for skip_rescale) */
  int out_act;         /* 0 none, 3 tanh */
  float* out;
  int out_mode;        /* 0 PNHWC, 1 NHWC, 2 NCHW */
  int out_C;           /* channel pitch of out/res for NHWC modes (0 = Cout) */
  double* stats;       /* [N][Cout][2] running sum / sum of squares of the stored values, or NULL */
  int precision;       /* 3 = BF16x3 split (fp32 parity), 1 = plain BF16 */
  int msub;            /* 0 auto, 1 or 2 accumulators of 128 rows per CTA */
  int force_linear;    /* 1: keep the 1-D padded-linear M tiling even where the 2-D (16 x 8) tiling applies (testing) */
  void* debug_prof;    /* optional int64[16] device buffer: per-role cycle counters of one CTA (tuning aid), or NULL */
  int batch_rows;      /* >0: batched GEMM (1x1 only): rows [b*batch_rows, (b+1)*batch_rows) use packed operand b
                          (wpack + b * ddg_conv_packed_bytes(...)); used for the attention GEMMs (layerspp.py:115-119) */
  int zero_border;     /* 1 (PNHWC output only): `out` is a fresh, uninitialised buffer -- its one-pixel frame is cleared first
                          (ddg_zero_border on the same stream); 0: the frame is already zero and is left alone */
  void* out_planes;    /* optional (PNHWC output only): also write the result as pre-split bf16 planes [planes][N][out_C/8][H+2][W+2][8]
                          for consumers that read it without a prologue (see ddg_conv_src.planes) */
  void* splitk_ws;     /* optional split-K workspace (device memory, zero-initialised once by the caller; any number of launches on ONE
                          stream may share it).  With it, layers whose grid would leave most SMs idle (4x4 / 8x8 levels) run 2 or 4
                          CTAs per output tile, each over a share of K; the partial sums meet in this buffer.  NULL: never split. */
  long splitk_ws_bytes;/* size of splitk_ws; layers that would need more fall back to one CTA per tile */
} ddg_conv_desc;

/* output-channel tile width the conv kernel will use for a problem with m_rows GEMM rows (pack and launch must agree) */
int ddg_conv_tile_n(int cout, long m_rows);
/* tuning switch: use N = 256 output-channel tiles when Cout % 256 == 0 (returns the previous setting) */
int ddg_conv_set_nt256(int on);
long ddg_conv_packed_bytes(int cout, int total_stages, int kb, int precision, int nt);
/* w[co*s_co + ci*s_ci + tap*s_tap] -> packed stages [stage_offset, stage_offset + cin_pad/kb*ntaps) of every n-tile */
/* batch > 1: `batch` operands, w advances by w_batch_stride floats, out by ddg_conv_packed_bytes(...) bytes per batch */
int ddg_conv_pack_weights(const float* w, void* out, int cout, int cin_real, int cin_pad, int ntaps, long s_co, long s_ci,
                          long s_tap, int flip_taps, int kb, int stage_offset, int total_stages, int precision, int nt, int batch,
                          long w_batch_stride, cudaStream_t stream);
/* The same packing for MANY operands in one launch: `items_dev` is a device-resident table (built once per network: the
 * weights live at fixed addresses), item i covering chunks [chunk_begin, chunk_begin + ddg_conv_pack_chunks(...)) of the
 * launch; total_chunks = the sum.  Per-item fields mean what the ddg_conv_pack_weights arguments mean (batch = 1). */
typedef struct {
  const float* w; void* out;
  long s_co, s_ci, s_tap;
  long chunk_begin;
  int cout, cin_real, cin_pad, ntaps, flip_taps, kb, stage_offset, total_stages, precision, nt;
} ddg_pack_item;
long ddg_conv_pack_chunks(int cout, int cin_pad, int ntaps, int kb, int nt);
int ddg_conv_pack_batch(const ddg_pack_item* items_dev, int n_items, long total_chunks, cudaStream_t stream);
int ddg_conv2d_fwd(const ddg_conv_desc* desc, cudaStream_t stream);
/* x: PNHWC fp32 [N][H+2][W+2][C] -> pre-split operand planes (layout above; `planes` = 2 for precision 3, 1 for precision 1);
 * only the interior is written (the planes buffer is zero-initialised once by its owner). */
int ddg_split_planes(const float* x, void* planes, int N, int H, int W, int C, int nplanes, cudaStream_t stream);
/* bytes of one plane of such a buffer */
long ddg_planes_bytes(int N, int H, int W, int C);
/* diagnostics: which kernel variant the last ddg_conv2d_fwd call on this host thread launched (tests assert on it) */
int ddg_conv_last_launch_info(int* msub, int* nt, int* persistent, int* grid_ctas);
/* ... and how many of its K segments were fetched by the TMA engine from pre-split planes */
int ddg_conv_last_launch_tma(void);
int ddg_conv_last_launch_ksplit(void);   /* split-K factor of the last ddg_conv2d_fwd launch made from this thread (1 = not split) */

/* Fused attention core of AttnBlockpp (layerspp.py:108-124) for 256 tokens x 256 channels (the 16x16 attention level):
 *   out = (res + NIN_3(softmax(q k^T / sqrt(C)) v) + bias) * out_scale, written PNHWC, + GroupNorm statistics of the result.
 * qkv: fp32 [N][T][3C] = q | k | v per token (the NHWC output of the fused q/k/v 1x1 conv); w3pack: NIN_3 weights packed by
 * ddg_conv_pack_weights with nt = 256 (one segment of C channels, 1 tap).  One CTA per (sample, 128 queries): logits, softmax
 * weights and the attention output stay in TMEM / shared memory. */
typedef struct {
  const float* qkv;
  const void* w3pack;
  const float* bias;   /* [C] or NULL */
  const float* res;    /* PNHWC [N][H+2][W+2][C] or NULL */
  float* out;          /* PNHWC */
  double* stats;       /* [N][C][2] or NULL */
  int N, H, W, C;
  float out_scale;
  int precision;       /* 3 = BF16x3, 1 = BF16 */
} ddg_attn_desc;
int ddg_attention_fwd(const ddg_attn_desc* desc, cudaStream_t stream);

/* Weight gradient of the same convolution (cuDNN wgrad in the reference's backward, ddgan.py:459-506):
 *   dw[co*s_co + ci*s_ci + tap*s_tap] += sum_q dy[q][co] * x[q + tap_dr*Wp + tap_ds][ci]   over the padded space [N][Hp][Wp]
 * x: PNHWC source as seen by the conv (pitch xpitch, Cin_pad channels used, Cin_real scattered);
 * dy: PNHWC output gradient with zero border (pitch dypitch, dy_cpad channels present).  dw must be zero-initialised
 * by the caller (the kernel accumulates with fp32 reductions across its split-K CTAs). */
typedef struct {
  const float* x; const float* dy; float* dw;
  int xpitch, dypitch;
  int N, Hp, Wp;
  int Cout, dy_cpad, Cin_real, Cin_pad;
  int ntaps;
  int8_t tap_dr[9]; int8_t tap_ds[9];
  long s_co, s_ci, s_tap;
  int precision;
  float gain;                  /* dw += gain * (...); 0 is read as 1 (the adjoint of a conv whose epilogue applies out_scale) */
  void* debug_prof;            /* optional int64[8] device buffer: per-role cycle counters of one CTA (tuning aid), or NULL */
} ddg_wgrad_desc;
int ddg_conv2d_wgrad(const ddg_wgrad_desc* desc, cudaStream_t stream);

/* ---- PNHWC training helpers --------------------------------------------------------------------------------------
 * y = act(scale[n,c] * x + shift[n,c]) on the interior: AdaGN / GN apply + SiLU (layerspp.py:279,300).  ddg_affine_act_fwd and
 * ddg_gn_bwd_dx also clear the one-pixel frame of their output (it may be a fresh, uninitialised buffer). */
int ddg_affine_act_fwd(const float* x, const float* scale, const float* shift, float* y, int N, int H, int W, int C, int act,
                       cudaStream_t stream);
/* dx = dy * act'(u) * scale, sums[n][c] = {sum dy*act'(u)*x, sum dy*act'(u)} (float64, accumulated: zero it first) */
int ddg_affine_act_bwd(const float* x, const float* dy, const float* scale, const float* shift, float* dx, double* sums, int N,
                       int H, int W, int C, int act, cudaStream_t stream);
/* GroupNorm backward in two passes over the activation (layerspp.py:46-63 under autograd in the reference):
 *   pass 1 = ddg_affine_act_bwd with dx = NULL: sums[n][c] = {sum gg*x, sum gg}, gg = dy * act'(scale*x + shift)
 *   ddg_gn_bwd_coeffs: from the forward statistics and those sums, d(gamma), d(beta) per (n, c) (written with row pitch
 *            dgb_stride, e.g. straight into a slice of the batched style-projection gradient) and g12[n][c] = {d(sum x), d(sum x^2)}
 *   pass 2 = ddg_gn_bwd_dx: dx = gg * scale + g12[n][c][0] + 2 * x * g12[n][c][1]. */
int ddg_gn_bwd_coeffs(const double* stats, const double* sums, const float* gamma, int gb_stride, int per_sample, float* g12,
                      float* dgamma, float* dbeta, int dgb_stride, int N, int C, int HW, int G, float eps, cudaStream_t stream);
int ddg_gn_bwd_dx(const float* x, const float* dy, const float* scale, const float* shift, const float* g12, float* dx, int N, int H,
                  int W, int C, int act, cudaStream_t stream);
/* stats[n][c] = {sum x, sum x^2} over the interior (float64, accumulated: zero it first) */
int ddg_stats_fwd(const float* x, double* stats, int N, int H, int W, int C, cudaStream_t stream);
/* dx = g1[n][c] + 2 * x * g2[n][c] on the interior, g = [N][C][2] float32 */
int ddg_stats_bwd(const float* x, const float* g, float* dx, int N, int H, int W, int C, cudaStream_t stream);

/* Bias and per-(sample, channel) gradients of a conv from its PNHWC output gradient dy [N][H+2][W+2][C] (C % 32 == 0):
 *   dav[n*dav_stride + c] = scale * sum_{h,w} dy[n][h][w][c]   (written; when ddg_channel_grads_splits(...) > 1 it is
 *                                                                accumulated instead and must be zeroed by the caller)
 *   db[c]               += scale * sum_{n,h,w} dy[n][h][w][c]   (accumulated: bias.grad or a zeroed buffer)
 * for c < cout; either output may be NULL.  The gradients of layerspp.py:298-299 (Dense_0) and of every conv bias. */
int ddg_channel_grads_splits(int N, int H, int W, int C);
int ddg_channel_grads(const float* dy, float* dav, float* db, int N, int H, int W, int C, int cout, float scale, int dav_stride,
                      cudaStream_t stream);
/* conv_downsample_2d weights (up_or_down_sampling.py:149-183): wt [Cout][Cin][3][3] -> w2 [Cout][2][2][cp][2][2] with
 * w2[co][py][px][ci][dy][dx] = wt[co][ci][2dy+py][2dx+px] (zero outside the 3x3 support / ci >= Cin): the stride-2 conv after the
 * pad-(2,2) FIR as a stride-1 2x2-tap conv over space-to-depth channels.  adjoint=1: src = d(w2), dst = d(wt). */
int ddg_s2d_weights(const float* src, float* dst, int Cout, int Cin, int cp, int adjoint, cudaStream_t stream);

/* Image output (test_ddgan.py:190-201): out[n][h][w][c] = (uint8) clamp((x[n][c][h][w]*scale + shift)*255 + 0.5, 0, 255), i.e.
 * to_range_0_1 (scale = shift = 0.5) followed by torchvision.utils.save_image's quantisation, for the whole NCHW batch at once. */
int ddg_images_to_u8(const float* x, uint8_t* out, int N, int C, int H, int W, float scale, float shift, cudaStream_t stream);

/* ---- flat-arena optimiser pass (ddgan.py:484-485, 507-508 clip_grad_norm_ + Adam; ema.py:45-55) --------------------------
 * out[0] = sum of squares of the gradient arena (float64; zeroed by the call). */
int ddg_grad_norm_sq(const float* g, long n, double* out, cudaStream_t stream);
/* One Adam step (torch.optim.Adam semantics, no amsgrad) over flat arenas with the gradient clipped to max_norm using
 * *normsq (NULL or max_norm <= 0: no clipping) and, when ema != NULL, ema = decay*ema + (1-decay)*p_new.
 * state: device float[2] = {step count (incremented by the call), learning rate}.
 * grad_scale multiplies the gradient (and its norm) first: 1/world_size after a sum all-reduce of the arena (ddgan.py:363-365
 * DistributedDataParallel averages), 1 otherwise. */
int ddg_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, long n, float* state, const double* normsq,
                      float max_norm, float beta1, float beta2, float eps, float weight_decay, float ema_decay,
                      float grad_scale, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
