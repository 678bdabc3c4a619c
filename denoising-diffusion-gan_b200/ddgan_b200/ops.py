"""Thin, allocation-explicit Python wrappers over the C ABI (one function per entry point).

Nothing here computes on the host or falls back to torch ops: each wrapper validates, allocates the output with
torch (device memory plumbing only) and calls libddgan_b200.so on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import math

import os
import torch

from . import _lib
from ._lib import ConvDesc, check, lib, ptr, require_cuda_f32, stream

ACT_NONE, ACT_SILU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3
OUT_PNHWC, OUT_NHWC, OUT_NCHW = 0, 1, 2
KB = 32

TAPS_3X3 = [(r - 1, s - 1) for r in range(3) for s in range(3)]
TAPS_2X2 = [(dy, dx) for dy in range(2) for dx in range(2)]
TAPS_1X1 = [(0, 0)]


def pad_c(c: int, m: int = KB) -> int:
    return (c + m - 1) // m * m


# ---------------------------------------------------------------------------------------------------------
# score_sde.op surface primitives
# ---------------------------------------------------------------------------------------------------------
def upfirdn2d_raw(x: torch.Tensor, k: torch.Tensor, up_x, up_y, down_x, down_y, px0, px1, py0, py1) -> torch.Tensor:
    """x [planes, H, W] -> [planes, H', W'] (upfirdn2d.cpp:20-31 argument order).  fp16 / bf16 inputs go to the 16-bit kernels
    (same up / down / pad on both axes, as the Python surface always passes)."""
    if x.dtype in _lib.LP_DTYPES:
        if not x.is_cuda:
            raise RuntimeError('ddgan_b200 ops run on CUDA tensors only (no CPU fallback)')
        if not (up_x == up_y and down_x == down_y and px0 == py0 and px1 == py1):
            raise NotImplementedError('16-bit upfirdn2d: up / down / pad must be the same on both axes')
        x = x.contiguous()
        k = k.to(torch.float32).contiguous()
        planes, in_h, in_w = x.shape
        kh, kw = k.shape
        out = torch.empty(planes, (in_h * up_y + py0 + py1 - kh) // down_y + 1, (in_w * up_x + px0 + px1 - kw) // down_x + 1,
                          device=x.device, dtype=x.dtype)
        check(lib().ddg_upfirdn2d_lp(ptr(x), ptr(k), ptr(out), planes, in_h, in_w, kh, kw, up_x, down_x, px0, px1,
                                     _lib.LP_DTYPES[x.dtype], stream()), 'upfirdn2d_lp')
        return out
    require_cuda_f32(x, k)
    x = x.contiguous()
    k = k.contiguous()
    planes, in_h, in_w = x.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (in_w * up_x + px0 + px1 - kw) // down_x + 1
    out = torch.empty(planes, out_h, out_w, device=x.device, dtype=torch.float32)
    check(lib().ddg_upfirdn2d(ptr(x), ptr(k), ptr(out), planes, in_h, in_w, kh, kw, up_x, up_y, down_x, down_y,
                              px0, px1, py0, py1, stream()), 'upfirdn2d')
    return out


def fused_bias_act(x, b, ref, act: int, grad: int, alpha: float, scale: float) -> torch.Tensor:
    """fused_bias_act.cpp:18-28; b / ref may be None or 0-element ('absent').  fp16 / bf16 x (and ref) use the 16-bit kernel."""
    if x.dtype in _lib.LP_DTYPES:
        if not x.is_cuda:
            raise RuntimeError('ddgan_b200 ops run on CUDA tensors only (no CPU fallback)')
        x = x.contiguous()
        b = None if (b is None or b.numel() == 0) else b.to(torch.float32).contiguous()
        ref = None if (ref is None or ref.numel() == 0) else ref.to(x.dtype).contiguous()
        y = torch.empty_like(x)
        step_b = 1
        for d in x.shape[2:]:
            step_b *= d
        check(lib().ddg_fused_bias_act_lp(ptr(x), ptr(b), ptr(ref), ptr(y), x.numel(), step_b, b.numel() if b is not None else 1, act,
                                          grad, alpha, scale, _lib.LP_DTYPES[x.dtype], stream()), 'fused_bias_act_lp')
        return y
    require_cuda_f32(x)
    x = x.contiguous()
    b = None if (b is None or b.numel() == 0) else b.contiguous()
    ref = None if (ref is None or ref.numel() == 0) else ref.contiguous()
    y = torch.empty_like(x)
    step_b = 1
    for d in x.shape[2:]:
        step_b *= d
    size_b = b.numel() if b is not None else 1
    check(lib().ddg_fused_bias_act(ptr(x), ptr(b), ptr(ref), ptr(y), x.numel(), step_b, size_b, act, grad, alpha, scale,
                                   stream()), 'fused_bias_act')
    return y


def channel_sum(g: torch.Tensor) -> torch.Tensor:
    """[N, C, ...] -> [C]"""
    require_cuda_f32(g)
    g = g.contiguous()
    n, c = g.shape[0], g.shape[1]
    inner = g.numel() // max(n * c, 1)
    out = torch.empty(c, device=g.device, dtype=torch.float32)
    check(lib().ddg_channel_sum(ptr(g), ptr(out), n, c, inner, stream()), 'channel_sum')
    return out


def groupnorm_fwd(x, G, gamma=None, beta=None, eps=1e-6, per_sample=False, act=ACT_NONE):
    """NCHW GroupNorm(+AdaGN affine, + activation). Returns (y, mean[N*G], rstd[N*G])."""
    require_cuda_f32(x, gamma, beta)
    x = x.contiguous()
    n, c = x.shape[:2]
    hw = x.numel() // (n * c)
    y = torch.empty_like(x)
    mean = torch.empty(n * G, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    if gamma is not None:
        gamma, beta = gamma.contiguous(), beta.contiguous()
    check(lib().ddg_groupnorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), n, c, hw, G, eps,
                                  int(per_sample), act, stream()), 'groupnorm_fwd')
    return y, mean, rstd


def groupnorm_bwd(x, dy, G, mean, rstd, gamma=None, beta=None, per_sample=False, act=ACT_NONE, need_affine_grads=True):
    require_cuda_f32(x, dy, gamma, beta, mean, rstd)
    x, dy = x.contiguous(), dy.contiguous()
    n, c = x.shape[:2]
    hw = x.numel() // (n * c)
    dx = torch.empty_like(x)
    dg = torch.empty(n, c, device=x.device, dtype=torch.float32) if need_affine_grads else None
    db = torch.empty(n, c, device=x.device, dtype=torch.float32) if need_affine_grads else None
    check(lib().ddg_groupnorm_bwd(ptr(x), ptr(dy), ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(dx), ptr(dg), ptr(db),
                                  n, c, hw, G, int(per_sample), act, stream()), 'groupnorm_bwd')
    return dx, dg, db


# ---------------------------------------------------------------------------------------------------------
# small fused kernels
# ---------------------------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int, max_positions: float = 10000.0, out=None) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.int64
    if out is None:
        out = torch.empty(t.shape[0], dim, device=t.device, dtype=torch.float32)
    check(lib().ddg_timestep_embedding(ptr(t), ptr(out), t.shape[0], dim, float(max_positions), stream()), 'timestep_embedding')
    return out


def linear(x, W, b=None, act_in=ACT_NONE, act_out=ACT_NONE, pixel_norm=False, out=None):
    """y = act_out(act_in(x) @ W.T + b); x [N, K] (row pitch = stride(0)), W [J, K]."""
    require_cuda_f32(x, W, b)
    assert x.stride(1) == 1 and W.is_contiguous()
    n, k = x.shape
    j = W.shape[0]
    if out is None:
        out = torch.empty(n, j, device=x.device, dtype=torch.float32)
    assert out.stride(1) == 1
    check(lib().ddg_linear(ptr(x), ptr(W), ptr(b), ptr(out), n, k, j, x.stride(0), out.stride(0), act_in, act_out,
                           int(pixel_norm), stream()), 'linear')
    return out


def make_mlp_desc(weights, biases, act=ACT_SILU, pixel_norm=False):
    """Descriptor of ddg_mlp_rows for nn.Linear-layout layers (W_i [out, in]); keeps no reference to the tensors."""
    d = _lib.MlpDesc()
    assert 1 <= len(weights) <= _lib.MLP_MAX_LAYERS
    for i, (w, b) in enumerate(zip(weights, biases)):
        require_cuda_f32(w, b)
        assert w.is_contiguous() and (b is None or b.is_contiguous())
        d.W[i] = w.data_ptr(); d.b[i] = b.data_ptr() if b is not None else None
        d.dims[i] = w.shape[1]; d.dims[i + 1] = w.shape[0]
        if i > 0:
            assert weights[i - 1].shape[0] == w.shape[1], 'layer widths do not chain'
    d.nlayers = len(weights); d.pixel_norm = int(pixel_norm); d.act = act
    return d


def mlp_rows(x, desc, out):
    """out[N, J] = the whole MLP described by `desc` applied to the rows of x (one kernel launch)."""
    require_cuda_f32(x, out)
    assert x.stride(1) == 1 and out.stride(1) == 1 and x.shape[1] == desc.dims[0] and out.shape[1] == desc.dims[desc.nlayers]
    check(lib().ddg_mlp_rows(ptr(x), x.stride(0), ptr(out), out.stride(0), x.shape[0], C.byref(desc), stream()), 'mlp_rows')
    return out


def q_sample_pairs(x0, noise_xt, noise_xtp1, t, a_s_cum, sigmas_cum, a_s, sigmas, out=None):
    require_cuda_f32(x0, noise_xt, noise_xtp1, a_s_cum, sigmas_cum, a_s, sigmas)
    x0, noise_xt, noise_xtp1 = x0.contiguous(), noise_xt.contiguous(), noise_xtp1.contiguous()
    x_t, x_tp1 = out if out is not None else (torch.empty_like(x0), torch.empty_like(x0))
    n = x0.shape[0]
    check(lib().ddg_q_sample_pairs(ptr(x0), ptr(noise_xt), ptr(noise_xtp1), ptr(t), ptr(a_s_cum), ptr(sigmas_cum), ptr(a_s),
                                   ptr(sigmas), ptr(x_t), ptr(x_tp1), n, x0.numel() // n, stream()), 'q_sample_pairs')
    return x_t, x_tp1


def sample_posterior(x0, x_t, noise, t, coef1, coef2, logvar, out=None):
    require_cuda_f32(x0, x_t, noise, coef1, coef2, logvar)
    x0, x_t, noise = x0.contiguous(), x_t.contiguous(), noise.contiguous()
    if out is None:
        out = torch.empty_like(x_t)
    n = x_t.shape[0]
    check(lib().ddg_sample_posterior(ptr(x0), ptr(x_t), ptr(noise), ptr(t), ptr(coef1), ptr(coef2), ptr(logvar), ptr(out), n,
                                     x_t.numel() // n, stream()), 'sample_posterior')
    return out


# ---------------------------------------------------------------------------------------------------------
# internal layout (PNHWC) helpers
# ---------------------------------------------------------------------------------------------------------
_POISON = bool(os.environ.get('DDG_POISON_ALLOC'))   # test aid: NaN-fill fresh interiors to catch kernels that skip elements


class FramePool:
    """Per-training-step pool of padded NHWC buffers whose one-pixel frame stays zero for good.

    Inside train.Trainer.step the convolution / FIR outputs of the differentiable graphs are taken from here in call order (one
    cursor per shape, rewound at the start of every step): the buffer a layer used in step s serves the same call in step s + 1.  Its
    producers write the interior only, nothing ever writes the frame, so the frame is cleared once at allocation instead of by one
    ddg_zero_border launch per tensor per step (~270 launches, 0.75 ms of the CIFAR-10 train step).  Buffers of one step are distinct
    (autograd keeps them alive until the backward pass has used them); the next step starts after everything that read them has been
    enqueued on the same stream.  Never active outside a Trainer step: the engines' static buffers and direct module calls allocate
    normally."""

    def __init__(self):
        self.bufs, self.cur, self.active = {}, {}, False

    def begin_step(self):
        self.cur = dict.fromkeys(self.bufs, 0)
        self.active = True

    def end_step(self):
        self.active = False

    def get(self, key, make):
        lst = self.bufs.setdefault(key, [])
        i = self.cur.get(key, 0)
        if i == len(lst):
            lst.append(make())
        self.cur[key] = i + 1
        # a fresh tensor object over the same storage every time: the object handed out last step carries that step's autograd
        # history and hook table (a hook registered on it again would never reach the new grad_fn: Tensor.register_hook attaches
        # the table to the node only when it creates it -- seen as the data-parallel early all-reduce not firing from step 2 on)
        return lst[i].detach()

    def clear(self):
        self.bufs, self.cur = {}, {}


FRAME_POOL = FramePool()


def alloc_pnhwc(n, h, w, c, device, full=True, border=True, pooled=False) -> torch.Tensor:
    """[N, H+2, W+2, C] with a zero border.  full=False: only the frame is cleared (for producers that write every interior
    element and channel); full=True zero-fills the whole buffer; border=False: nothing is cleared -- the producing kernel
    writes the frame itself (ddg_affine_act_fwd, ddg_gn_bwd_dx).  pooled=True (training graphs): inside a Trainer step the
    buffer comes from FRAME_POOL (see there) when only the frame has to be zero."""
    if pooled and border and not full and c % 4 == 0 and FRAME_POOL.active and not _POISON:
        dev = torch.device(device)
        return FRAME_POOL.get((dev.type, dev.index, n, h, w, c), lambda: alloc_pnhwc(n, h, w, c, device, full=False, border=True))
    if full or c % 4 != 0:
        return torch.zeros(n, h + 2, w + 2, c, device=device, dtype=torch.float32)
    out = torch.empty(n, h + 2, w + 2, c, device=device, dtype=torch.float32)
    if _POISON:
        out.fill_(float('nan'))
    if border:
        check(lib().ddg_zero_border(ptr(out), n, h, w, c, stream()), 'zero_border')
    return out


def empty_like_pnhwc(x, border=True) -> torch.Tensor:
    n, hp, wp, c = x.shape
    return alloc_pnhwc(n, hp - 2, wp - 2, c, x.device, full=False, border=border)


def to_pnhwc(a, b=None, cpad=None, out=None, scale=1.0, shift=0.0):
    require_cuda_f32(a, b)
    a = a.contiguous()
    n, ca, h, w = a.shape
    cb = 0
    if b is not None:
        b = b.contiguous()
        cb = b.shape[1]
    if cpad is None:
        cpad = pad_c(ca + cb)
    if out is None:
        out = alloc_pnhwc(n, h, w, cpad, a.device, full=False)
    check(lib().ddg_nchw_to_pnhwc(ptr(a), ca, ptr(b), cb, ptr(out), n, h, w, cpad, scale, shift, stream()), 'nchw_to_pnhwc')
    return out


def from_pnhwc(x, c=None, padded=True, out=None):
    require_cuda_f32(x)
    if padded:
        n, hp, wp, cp = x.shape
        h, w = hp - 2, wp - 2
    else:
        n, h, w, cp = x.shape
    c = c or cp
    if out is None:
        out = torch.empty(n, c, h, w, device=x.device, dtype=torch.float32)
    check(lib().ddg_pnhwc_to_nchw(ptr(x), ptr(out), n, h, w, c, cp, int(padded), stream()), 'pnhwc_to_nchw')
    return out


def gn_prepare(stats_a, ca, stats_b, cb, gamma, beta, gb_stride, per_sample, n, hw, groups, scale, shift, eps=1e-6):
    check(lib().ddg_gn_prepare(ptr(stats_a), ca, ptr(stats_b), cb, ptr(gamma), ptr(beta), gb_stride, int(per_sample),
                               ptr(scale), ptr(shift), n, hw, groups, eps, stream()), 'gn_prepare')


def gn_prepare_bwd(stats, gamma, gb_stride, per_sample, dscale, dshift, n, c, hw, groups, eps=1e-6):
    dstats = torch.empty(n, c, 2, dtype=torch.float64, device=stats.device)
    dgamma = torch.empty(n, c, dtype=torch.float32, device=stats.device)
    dbeta = torch.empty(n, c, dtype=torch.float32, device=stats.device)
    check(lib().ddg_gn_prepare_bwd(ptr(stats), ptr(gamma), gb_stride, int(per_sample), ptr(dscale), ptr(dshift), ptr(dstats),
                                   ptr(dgamma), ptr(dbeta), n, c, hw, groups, eps, stream()), 'gn_prepare_bwd')
    return dstats, dgamma, dbeta


def fir_pnhwc(x, mode, out, scale=None, shift=None, act=ACT_NONE, gain=1.0):
    """mode 1 up2, 2 down2, 3 pad(2,2)+space-to-depth, 4 adjoint of 3 (x is the s2d tensor, out the image)."""
    if mode == 4:
        n, hp, wp, c = out.shape
        check(lib().ddg_fir_pnhwc(ptr(x), None, None, ACT_NONE, ptr(out), n, hp - 2, wp - 2, c, 4, x.shape[-1], gain, stream()),
              'fir_pnhwc')
        return out
    n, hp, wp, c = x.shape
    check(lib().ddg_fir_pnhwc(ptr(x), ptr(scale), ptr(shift), act, ptr(out), n, hp - 2, wp - 2, c, mode, out.shape[-1], gain,
                              stream()), 'fir_pnhwc')
    return out


def affine_act_fwd(x, scale, shift, act, out=None):
    n, hp, wp, c = x.shape
    if out is None:
        out = empty_like_pnhwc(x, border=False)        # the kernel clears the frame itself
    check(lib().ddg_affine_act_fwd(ptr(x), ptr(scale), ptr(shift), ptr(out), n, hp - 2, wp - 2, c, act, stream()), 'affine_act_fwd')
    return out


def affine_act_bwd(x, dy, scale, shift, act, need_sums=True, need_dx=True, sums=None):
    """sums (float64 [N, C, 2], accumulated: pass a zeroed buffer or let one be allocated) = {sum gg*x, sum gg} with
    gg = dy * act'(scale*x + shift); dx = gg * scale unless need_dx is False (first pass of the fused GroupNorm backward)."""
    n, hp, wp, c = x.shape
    dx = empty_like_pnhwc(x) if need_dx else None
    if sums is None and need_sums:
        sums = torch.zeros(n, c, 2, dtype=torch.float64, device=x.device)
    check(lib().ddg_affine_act_bwd(ptr(x), ptr(dy), ptr(scale), ptr(shift), ptr(dx), ptr(sums), n, hp - 2, wp - 2, c, act, stream()),
          'affine_act_bwd')
    return dx, sums


def gn_bwd_coeffs(stats, sums, gamma, gb_stride, per_sample, g12, dgamma, dbeta, dgb_stride, n, c, hw, groups, eps=1e-6):
    check(lib().ddg_gn_bwd_coeffs(ptr(stats), ptr(sums), ptr(gamma), gb_stride, int(per_sample), ptr(g12), ptr(dgamma), ptr(dbeta),
                                  dgb_stride, n, c, hw, groups, eps, stream()), 'gn_bwd_coeffs')


def gn_bwd_dx(x, dy, scale, shift, g12, act):
    n, hp, wp, c = x.shape
    dx = empty_like_pnhwc(x, border=False)             # the kernel clears the frame itself
    check(lib().ddg_gn_bwd_dx(ptr(x), ptr(dy), ptr(scale), ptr(shift), ptr(g12), ptr(dx), n, hp - 2, wp - 2, c, act, stream()), 'gn_bwd_dx')
    return dx


def stats_fwd(x, out=None):
    n, hp, wp, c = x.shape
    st = out if out is not None else torch.zeros(n, c, 2, dtype=torch.float64, device=x.device)
    check(lib().ddg_stats_fwd(ptr(x), ptr(st), n, hp - 2, wp - 2, c, stream()), 'stats_fwd')
    return st


def stats_bwd(x, g):
    n, hp, wp, c = x.shape
    dx = empty_like_pnhwc(x)
    check(lib().ddg_stats_bwd(ptr(x), ptr(g), ptr(dx), n, hp - 2, wp - 2, c, stream()), 'stats_bwd')
    return dx


def conv_wgrad(x, dy, dw, n, hp, wp, cout, cin_real, cin_pad, taps, s_co, s_ci, s_tap, precision=3, xpitch=None, dypitch=None,
               dy_cpad=None, x_elem_offset=0, flops=None, gain=1.0, prof=None):
    """dw[co*s_co + ci*s_ci + t*s_tap] += sum_q dy[q][co] * x[q + tap_t][ci]  (dw zero-initialised by the caller)."""
    d = _lib.WgradDesc()
    d.x = x.data_ptr() + 4 * x_elem_offset; d.dy = ptr(dy); d.dw = ptr(dw)
    d.xpitch = xpitch or x.shape[-1]; d.dypitch = dypitch or dy.shape[-1]
    d.N, d.Hp, d.Wp = n, hp, wp
    d.Cout = cout; d.dy_cpad = dy_cpad or dy.shape[-1]; d.Cin_real = cin_real; d.Cin_pad = cin_pad
    d.ntaps = len(taps)
    for t, (dr, ds) in enumerate(taps):
        d.tap_dr[t] = dr; d.tap_ds[t] = ds
    d.s_co, d.s_ci, d.s_tap = s_co, s_ci, s_tap
    d.precision = precision
    d.gain = gain
    d.debug_prof = prof.data_ptr() if prof is not None else None
    if PROFILE['on']:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib().ddg_conv2d_wgrad(C.byref(d), stream()), 'conv2d_wgrad')
        b.record()
        PROFILE['records'].append(('wgrad_tc', flops if flops is not None else 2.0 * n * (hp - 2) * (wp - 2) * cout * cin_pad * len(taps), a, b))
        return dw
    check(lib().ddg_conv2d_wgrad(C.byref(d), stream()), 'conv2d_wgrad')
    return dw


def gemm_rows(x, W, b=None, precision=3):
    """y[N, J] = x[N, K] @ W[J, K]^T + b on the tensor-core conv kernel (a 1x1 conv over N 'pixels'); K % 32 == 0."""
    n, k = x.shape
    j = W.shape[0]
    cw = ConvWeights(j, [(k, 1)], x.device, precision=precision, m_rows=n)
    cw.pack_segment(0, W, k, k, 1, 0)
    out = torch.empty(n, j, device=x.device, dtype=torch.float32)
    d = build_conv_desc(cw, [conv_src(x, k, TAPS_1X1, padded=False)], n, 1, 1, out, out_mode=OUT_NHWC, bias=b)
    conv_launch(d)
    return out


def gemm_tn(a, b, precision=3):
    """out[Ca, Cb] = a[Q, Ca]^T @ b[Q, Cb] (contraction over rows) on the tensor-core weight-gradient kernel.
    Ca % 8 == 0, Cb % 32 == 0."""
    q, ca = a.shape
    cb = b.shape[1]
    out = torch.zeros(ca, cb, device=a.device, dtype=torch.float32)
    conv_wgrad(b, a, out, 1, q, 1, ca, cb, cb, [(0, 0)], cb, 1, 0, precision=precision, xpitch=cb, dypitch=ca, dy_cpad=ca,
               flops=2.0 * q * ca * cb)
    return out


def channel_grads(dy, cout, scale=1.0, need_dav=True, db_accum=None, need_db=True, dav_into=None):
    """(dav [N, cout] or None, db [cout] or None) = scale * sums of the PNHWC gradient dy over (h, w) / over (n, h, w).
    db_accum: accumulate the bias gradient straight into this tensor (a .grad view) instead of returning it.
    dav_into = (address, row pitch): write the per-sample sums into a slice of a wider zero-initialised buffer instead."""
    n, hp, wp, c = dy.shape
    dav = db = None
    dav_stride = cout
    if dav_into is not None:
        dav, dav_stride = dav_into
    elif need_dav:
        s = lib().ddg_channel_grads_splits(n, hp - 2, wp - 2, c)
        dav = (torch.zeros if s > 1 else torch.empty)(n, cout, device=dy.device, dtype=torch.float32)
    if db_accum is not None:
        dbp = db_accum
    elif need_db:
        db = torch.zeros(cout, device=dy.device, dtype=torch.float32)
        dbp = db
    else:
        dbp = None
    if dav is None and dbp is None:
        return None, None
    check(lib().ddg_channel_grads(ptr(dy), ptr(dav), ptr(dbp), n, hp - 2, wp - 2, c, cout, scale, dav_stride, stream()), 'channel_grads')
    return (None if dav_into is not None else dav), db


def s2d_weights(src, cout, cin, cp, adjoint=False, out=None):
    """conv_downsample_2d weights [Cout, Cin, 3, 3] <-> the 2x2-tap space-to-depth form [Cout, 2, 2, cp, 2, 2] (adjoint: gradient)."""
    require_cuda_f32(src)
    src = src.contiguous()
    if out is None:
        out = torch.empty((cout, cin, 3, 3) if adjoint else (cout, 2, 2, cp, 2, 2), device=src.device, dtype=torch.float32)
    check(lib().ddg_s2d_weights(ptr(src), ptr(out), cout, cin, cp, int(adjoint), stream()), 's2d_weights')
    return out


def minibatch_stddev(x, out, group):
    n, hp, wp, c = x.shape
    check(lib().ddg_minibatch_stddev(ptr(x), ptr(out), n, hp - 2, wp - 2, c, out.shape[-1], group, stream()), 'minibatch_stddev')
    return out


def spatial_sum(x, act=ACT_NONE, out=None):
    n, hp, wp, c = x.shape
    if out is None:
        out = torch.empty(n, c, device=x.device, dtype=torch.float32)
    check(lib().ddg_spatial_sum(ptr(x), ptr(out), n, hp - 2, wp - 2, c, act, stream()), 'spatial_sum')
    return out


def softmax_rows_bwd(p, dp, ds, rows, T, ld, scale):
    check(lib().ddg_softmax_rows_bwd(ptr(p), ptr(dp), ptr(ds), rows, T, ld, scale, stream()), 'softmax_rows_bwd')
    return ds


def bgemm(a, bsrc, cout, k_real, s_co, s_ci, b_batch_stride, precision=3, out_scale=1.0, out_c=0, elem_offset=0):
    """Batched GEMM on the tcgen05 conv kernel: out[n] = a[n] @ B[n]^T with a [N, M, Kp] fp32 row-major (Kp % 32 == 0) and
    B[n][co][ci] = bsrc.flat[elem_offset + n*b_batch_stride + co*s_co + ci*s_ci] for ci < k_real (zero beyond), packed per image.
    Returns [N, M, out_c or cout] (columns beyond cout are zero when out_c > cout)."""
    n, m, kp = a.shape
    assert a.is_contiguous() and kp % KB == 0
    cw = ConvWeights(cout, [(kp, 1)], a.device, precision=precision, batch=n, m_rows=n * m)
    cw.pack_segment(0, bsrc, k_real, s_co, s_ci, 0, w_batch_stride=b_batch_stride, elem_offset=elem_offset)
    oc = out_c or cout
    out = (torch.zeros if oc != cout else torch.empty)(n, m, oc, device=a.device, dtype=torch.float32)
    d = build_conv_desc(cw, [conv_src(a, kp, TAPS_1X1, padded=False)], n, 1, m, out, out_mode=OUT_NHWC, out_c=oc, out_scale=out_scale,
                        batch_rows=m)
    conv_launch(d)
    return out


def attention_desc(qkv, w3: 'ConvWeights', bias, res, out, stats, n, h, w, c, out_scale, precision=3):
    d = _lib.AttnDesc()
    d.qkv = ptr(qkv); d.w3pack = ptr(w3.buf); d.bias = _addr(bias); d.res = _addr(res); d.out = _addr(out); d.stats = _addr(stats)
    d.N, d.H, d.W, d.C = n, h, w, c
    d.out_scale = out_scale
    d.precision = precision
    assert w3.nt == 256 and w3.cout == c and tuple(w3.segs) == ((c, 1),)
    return d


def attention_launch(desc):
    """Fused attention core (ddg_attention_fwd): 256 tokens x 256 channels."""
    if PROFILE['on']:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib().ddg_attention_fwd(C.byref(desc), stream()), 'attention_fwd')
        b.record()
        PROFILE['records'].append(('attn_tc', 2.0 * desc.N * 256 * 256 * 256 * 3, a, b))
        return
    check(lib().ddg_attention_fwd(C.byref(desc), stream()), 'attention_fwd')


def softmax_rows(s, p, rows, T, lds, ldp):
    check(lib().ddg_softmax_rows(ptr(s), ptr(p), rows, T, lds, ldp, stream()), 'softmax_rows')
    return p


# ---------------------------------------------------------------------------------------------------------
# tcgen05 implicit-GEMM convolution
# ---------------------------------------------------------------------------------------------------------
class PackPlan:
    """All weight packs of a network as ONE launch (ddg_conv_pack_batch): a device-resident table of pack items, recorded the
    first time the packs are issued one by one and replayed afterwards.  Items point at fixed addresses (parameters in a flat
    arena or engine-owned copies, packed buffers owned by ConvWeights objects that the plan keeps alive)."""

    def __init__(self):
        self.items = []
        self.keep = []
        self.table = None
        self.total = 0

    def add(self, cw, seg, w_addr, cin_real, s_co, s_ci, s_tap, flip):
        c, nt = cw.segs[seg]
        it = _lib.PackItem()
        it.w = w_addr; it.out = cw.buf.data_ptr()
        it.s_co, it.s_ci, it.s_tap = s_co, s_ci, s_tap
        it.chunk_begin = self.total
        it.cout, it.cin_real, it.cin_pad, it.ntaps, it.flip_taps = cw.cout, cin_real, c, nt, int(flip)
        it.kb, it.stage_offset, it.total_stages, it.precision, it.nt = KB, cw.offsets[seg], cw.total_stages, cw.precision, cw.nt
        self.total += lib().ddg_conv_pack_chunks(cw.cout, c, nt, KB, cw.nt)
        self.items.append(it)
        self.keep.append(cw)
        self.table = None

    def finalize(self, device):
        if not self.items:
            return
        arr = (_lib.PackItem * len(self.items))(*self.items)
        raw = bytes(arr)
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)

    def run(self):
        if not self.items:
            return
        if self.table is None:
            self.finalize(self.keep[0].buf.device)
        check(lib().ddg_conv_pack_batch(ptr(self.table), len(self.items), self.total, stream()), 'conv_pack_batch')


# Pack recorder: while set to (plan, mode), ConvWeights.pack_segment does not launch; mode 'record' appends the pack to the plan,
# mode 'skip' assumes it is already in the plan (the caller runs plan.run() afterwards).
_PACK_CTX = [None]


class pack_context:
    def __init__(self, plan, mode):
        self.ctx = (plan, mode)

    def __enter__(self):
        self.prev = _PACK_CTX[0]
        _PACK_CTX[0] = self.ctx

    def __exit__(self, *a):
        _PACK_CTX[0] = self.prev


class ConvWeights:
    """Device-resident packed B operand of one fused convolution (all K segments, all n-tiles).

    segs: list of (C_padded, ntaps).  Call pack_segment() for each segment whenever the fp32 weights change."""

    def __init__(self, cout: int, segs, device, precision: int = 3, batch: int = 1, m_rows: int = 0, nt: int = 0):
        """m_rows: GEMM rows of the conv that will consume these weights (picks the output-channel tile width); nt forces it."""
        self.cout = cout
        self.nt = nt if nt else lib().ddg_conv_tile_n(cout, m_rows)
        self.segs = list(segs)
        self.precision = precision
        self.batch = batch
        self.total_stages = sum((c // KB) * nt for c, nt in self.segs)
        self.bytes_per_batch = lib().ddg_conv_packed_bytes(cout, self.total_stages, KB, precision, self.nt)
        self.buf = torch.empty(self.bytes_per_batch * batch, dtype=torch.uint8, device=device)
        self.offsets = []
        off = 0
        for c, nt in self.segs:
            self.offsets.append(off)
            off += (c // KB) * nt

    def pack_segment(self, i: int, w: torch.Tensor, cin_real: int, s_co: int, s_ci: int, s_tap: int, flip: bool = False,
                     w_batch_stride: int = 0, elem_offset: int = 0):
        """B[co][ci][tap] = w.flat[elem_offset + co*s_co + ci*s_ci + tap*s_tap] for ci < cin_real (zero beyond)."""
        require_cuda_f32(w)
        c, nt = self.segs[i]
        if _PACK_CTX[0] is not None and self.batch == 1:
            plan, mode = _PACK_CTX[0]
            if mode == 'record':
                plan.add(self, i, w.data_ptr() + 4 * elem_offset, cin_real, s_co, s_ci, s_tap, flip)
            return
        check(lib().ddg_conv_pack_weights(w.data_ptr() + 4 * elem_offset, ptr(self.buf), self.cout, cin_real, c, nt, s_co, s_ci, s_tap, int(flip),
                                          KB, self.offsets[i], self.total_stages, self.precision, self.nt, self.batch, w_batch_stride,
                                          stream()), 'conv_pack_weights')

    def pack_conv_weight(self, i: int, w: torch.Tensor):
        """w: [Cout, Cin, kh, kw] contiguous (nn.Conv2d layout)."""
        assert w.is_contiguous()
        co, ci, kh, kw = w.shape
        assert kh * kw == self.segs[i][1] and co == self.cout
        self.pack_segment(i, w, ci, ci * kh * kw, kh * kw, 1)

    def pack_nin_weight(self, i: int, W: torch.Tensor):
        """W: [in, out] contiguous (layers.py:492 NIN)."""
        assert W.is_contiguous() and W.shape[1] == self.cout
        self.pack_segment(i, W, W.shape[0], 1, W.shape[1], 0)


def conv_src(x, c, taps, scale=None, shift=None, act=ACT_NONE, padded=True, pitch=0, ss_stride=0, planes=None, planes_C=0, planes_c0=0):
    """x / scale / shift may be tensors or raw device addresses (ints) for channel-sliced views.  planes: optional pre-split bf16
    copy of the same tensor (alloc_planes layout) -- prologue-free sources only; the kernel then fetches this K segment by TMA."""
    return dict(x=x, C=c, taps=taps, scale=scale, shift=shift, act=act, padded=padded, pitch=pitch, ss_stride=ss_stride,
                planes=planes, planes_C=planes_C or c, planes_c0=planes_c0)


def alloc_planes(n, h, w, c, precision, device):
    """Zero-initialised pre-split operand planes of a PNHWC tensor: bf16 [planes][N][C/8][H+2][W+2][8] (planes = 2 for BF16x3)."""
    npl = 2 if precision == 3 else 1
    return torch.zeros(npl * lib().ddg_planes_bytes(n, h, w, c), dtype=torch.uint8, device=device)


def alloc_splitk_ws(device, nbytes=8 << 20):
    """Workspace for the split-K convolutions of the small spatial levels (ddg_conv_desc.splitk_ws): arrival counters + fp32 partial
    tiles.  One buffer serves every convolution launched on one stream."""
    return torch.zeros(nbytes, dtype=torch.uint8, device=device)


def split_planes(x, planes, precision):
    n, hp, wp, c = x.shape
    check(lib().ddg_split_planes(ptr(x), ptr(planes), n, hp - 2, wp - 2, c, 2 if precision == 3 else 1, stream()), 'split_planes')
    return planes


def _addr(v):
    return v if (v is None or isinstance(v, int)) else v.data_ptr()


def build_conv_desc(weights: ConvWeights, srcs, n, hout, wout, out, out_mode=OUT_PNHWC, hp=None, wp=None, bias=None, addvec=None,
                    addvec_stride=0, res=None, out_scale=1.0, out_act=ACT_NONE, out_c=0, stats=None, msub=0, batch_rows=0, prof=None, force_linear=0,
                    out_planes=None, zero_border=0, splitk_ws=None) -> ConvDesc:
    d = ConvDesc()
    d.nsrc = len(srcs)
    for i, s in enumerate(srcs):
        sd = d.src[i]
        sd.x = _addr(s['x']); sd.scale = _addr(s.get('scale')); sd.shift = _addr(s.get('shift'))
        sd.C = s['C']; sd.pitch = s.get('pitch', 0); sd.ss_stride = s.get('ss_stride', 0); sd.act = s.get('act', ACT_NONE); sd.ntaps = len(s['taps']); sd.padded = int(s.get('padded', True))
        for t, (dr, ds) in enumerate(s['taps']):
            sd.tap_dr[t] = dr
            sd.tap_ds[t] = ds
        sd.planes = _addr(s.get('planes')); sd.planes_C = s.get('planes_C', 0) or s['C']; sd.planes_c0 = s.get('planes_c0', 0)
        assert (s['C'], len(s['taps'])) == tuple(weights.segs[i]), 'source / packed-weight segment mismatch'
    d.wpack = ptr(weights.buf)
    d.kb = KB
    d.nt = weights.nt
    d.N, d.Hout, d.Wout = n, hout, wout
    d.Hp = hp if hp is not None else hout + 2
    d.Wp = wp if wp is not None else wout + 2
    d.Cout = weights.cout
    d.bias = _addr(bias); d.addvec = _addr(addvec); d.addvec_stride = addvec_stride
    d.res = _addr(res); d.out_scale = out_scale; d.out_act = out_act
    d.out = _addr(out); d.out_mode = out_mode; d.out_C = out_c
    d.stats = _addr(stats)
    d.precision = weights.precision
    d.msub = msub
    d.batch_rows = batch_rows
    d.debug_prof = _addr(prof)
    d.force_linear = force_linear
    d.out_planes = _addr(out_planes)
    d.zero_border = int(zero_border)
    if splitk_ws is not None:          # uint8 tensor, zero-initialised once (alloc_splitk_ws); shared by the launches of one stream
        d.splitk_ws = splitk_ws.data_ptr()
        d.splitk_ws_bytes = splitk_ws.numel()
    return d


# Optional per-launch timing of the tensor-core kernels (bench.py roofline): CUDA events on the launching stream.
PROFILE = {'on': False, 'records': []}


def _conv_flops(d: ConvDesc):
    k = sum(d.src[i].C * d.src[i].ntaps for i in range(d.nsrc))
    return 2.0 * d.N * d.Hout * d.Wout * d.Cout * k


def conv_launch(desc: ConvDesc):
    if PROFILE['on']:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib().ddg_conv2d_fwd(C.byref(desc), stream()), 'conv2d_fwd')
        b.record()
        PROFILE['records'].append(('conv_tc', _conv_flops(desc), a, b))
        return
    check(lib().ddg_conv2d_fwd(C.byref(desc), stream()), 'conv2d_fwd')


def conv_last_launch_info():
    """(msub, nt, persistent, CTAs) of the last conv launch made from this thread."""
    v = [C.c_int(0) for _ in range(4)]
    check(lib().ddg_conv_last_launch_info(*[C.byref(a) for a in v]), 'conv_last_launch_info')
    return tuple(a.value for a in v)


def conv2d_fused(weights, srcs, n, hout, wout, out, **kw):
    d = build_conv_desc(weights, srcs, n, hout, wout, out, **kw)
    conv_launch(d)
    return out
