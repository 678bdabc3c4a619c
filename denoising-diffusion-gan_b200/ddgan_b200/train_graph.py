"""Differentiable (training) forwards of NCSNpp / Discriminator_* built from autograd Functions over the sm_100a kernels.

The adversarial step (ddgan.py:443-518) needs first-order gradients of G and D and, every `lazy_reg` steps, the
double-backward of D for the R1 penalty (ddgan.py:462-467).  The heavy operators are closed under differentiation:

    ConvFn   (implicit-GEMM forward)      backward -> DgradFn (w.r.t. x), WgradFn (w.r.t. w)
    DgradFn  (conv with adjoint taps)     backward -> ConvFn  (w.r.t. dy), WgradFn (w.r.t. w)
    WgradFn  (tcgen05 weight gradient)    backward -> DgradFn (w.r.t. x),  ConvFn  (w.r.t. dy)
    FirFn    (up / down / pad FIR)        backward -> FirFn with the adjoint mode and gain (upfirdn2d.py:119-122 analogue)
    ToPnhwcFn <-> FromPnhwcFn             layout changes at the model boundary

so `torch.autograd.grad(..., create_graph=True)` through D just records more of the same kernels.  GroupNorm / AdaGN is
decomposed as  StatsFn (per-(n,c) sums)  ->  tiny [N,C] algebra in torch (float64)  ->  AffineActFn (scale/shift + SiLU),
which makes autograd assemble the exact GroupNorm gradient.  Elementwise glue on activations (residual add, 1/sqrt2, cat,
LeakyReLU, tanh) and the attention core stay torch ops in this first training path; inference uses the fully fused plan in
engine.py instead.  Activations are PNHWC ([N,H+2,W+2,C], zero border); every op here preserves the zero border.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import arch, ops

RSQRT2 = 1.0 / math.sqrt(2.0)


# ------------------------------------------------------------------------------------------------------------------
# layout boundary
# ------------------------------------------------------------------------------------------------------------------
class ToPnhwcFn(Function):
    @staticmethod
    def forward(ctx, x, cpad):
        ctx.c = x.shape[1]
        return ops.to_pnhwc(x.contiguous(), cpad=cpad)

    @staticmethod
    def backward(ctx, g):
        return FromPnhwcFn.apply(g, ctx.c), None


class FromPnhwcFn(Function):
    @staticmethod
    def forward(ctx, x, c):
        ctx.cpad = x.shape[-1]
        return ops.from_pnhwc(x.contiguous(), c)

    @staticmethod
    def backward(ctx, g):
        return ToPnhwcFn.apply(g, ctx.cpad), None


# ------------------------------------------------------------------------------------------------------------------
# convolution family
# ------------------------------------------------------------------------------------------------------------------
# Set by Trainer.step: weight / bias gradients of leaf parameters that already own a .grad buffer (views into the flat gradient
# arena of train.FlatAdam) are accumulated straight into it by the wgrad / channel-sum kernels, and autograd is handed None --
# no per-parameter zero-filled temporary and no AccumulateGrad add (about 900 launches per step).  Off by default: plain
# autograd semantics (torch.autograd.grad must not touch .grad).
ACCUM = {'on': False}


class TrainPacks:
    """Packed B operands (forward and transposed / dgrad layouts) of a module's leaf conv weights.  The first training step packs
    them one by one and records the packs; after freeze() every forward re-packs all of them with ONE launch (PackPlan)."""

    def __init__(self):
        self.plan = ops.PackPlan()
        self.cw = {}
        self.frozen = False

    def begin(self):
        if self.frozen:
            self.plan.run()

    def freeze(self, device):
        if not self.frozen:
            self.plan.finalize(device)
            self.frozen = True


def train_packs(mod):
    ptrs = tuple(p.data_ptr() for p in mod.parameters())
    ent = getattr(mod, '_train_packs', None)
    if ent is None or ent[0] != ptrs:
        ent = (ptrs, TrainPacks())
        mod._train_packs = ent
    return ent[1]


def freeze_packs(mod):
    ent = getattr(mod, '_train_packs', None)
    if ent is not None:
        ent[1].freeze(next(mod.parameters()).device)


_CUR_PACKS = [None]

# split-K workspace of the small-level convolutions (ddg_conv_desc.splitk_ws): every conv of the training graphs is launched on the
# stream the step runs on (eager: autograd replays the backward on the forward's stream; captured: the capture stream), so one buffer
# per (device, stream) is enough
_SPLITK_WS = {}


def _splitk_ws(dev):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = _SPLITK_WS[key] = ops.alloc_splitk_ws(dev)
    return ws


# module -> callable fired in the backward pass when the up path + head have been differentiated (see generator_forward).  Kept
# outside the module so that copy.deepcopy / pickling of the network never drags a Trainer along.
import weakref
GRAD_READY_HOOKS = weakref.WeakKeyDictionary()


def conv_spec(w_shape, taps, n, hout, wout, cpad_in, s_co, s_ci, s_tap, cout, cin, hp=None, wp=None, out_nchw=False, prec=3, out_scale=1.0):
    return SimpleNamespace(w_shape=tuple(w_shape), taps=list(taps), n=n, hout=hout, wout=wout, hp=hp or hout + 2,
                           wp=wp or wout + 2, cpad_in=cpad_in, s_co=s_co, s_ci=s_ci, s_tap=s_tap, cout=cout, cin=cin,
                           out_nchw=out_nchw, prec=prec, cpad_out=ops.pad_c(cout), out_scale=float(out_scale), packs=_CUR_PACKS[0])


def _packed(w, w_ref, kind, sp, device, cy=None):
    """ConvWeights holding w as the B operand of the forward conv ('fwd') or, with co / ci swapped, of its dgrad ('dgrad')."""
    ntaps = len(sp.taps)
    if kind == 'fwd':
        cout, seg_c, cin_real, s_co, s_ci = sp.cout, sp.cpad_in, sp.cin, sp.s_co, sp.s_ci
        m_rows = sp.n * sp.hp * sp.wp if ntaps > 1 else sp.n * sp.hout * sp.wout
    else:
        cout, seg_c, cin_real, s_co, s_ci = sp.cpad_in, cy, sp.cout, sp.s_ci, sp.s_co
        m_rows = sp.n * sp.hp * sp.wp if ntaps > 1 else sp.n * (sp.hp - 2) * (sp.wp - 2)
    packs = sp.packs
    cacheable = packs is not None and isinstance(w_ref, torch.nn.Parameter) and w_ref.data_ptr() == w.data_ptr()
    if cacheable:
        key = (w.data_ptr(), kind, cout, seg_c, ntaps, cin_real, s_co, s_ci, sp.s_tap, sp.prec, m_rows)
        cw = packs.cw.get(key)
        if cw is not None:
            if not packs.frozen:
                cw.pack_segment(0, w, cin_real, s_co, s_ci, sp.s_tap)
            return cw                                     # frozen: packed by packs.begin() of this forward
    cw = ops.ConvWeights(cout, [(seg_c, ntaps)], device, precision=sp.prec, m_rows=m_rows)
    cw.pack_segment(0, w, cin_real, s_co, s_ci, sp.s_tap)
    if cacheable and not packs.frozen:
        packs.cw[key] = cw
        packs.plan.add(cw, 0, w.data_ptr(), cin_real, s_co, s_ci, sp.s_tap, False)
    return cw


def _conv_forward(x, w, bias, addvec, sp, res=None, w_ref=None):
    cw = _packed(w, w_ref, 'fwd', sp, x.device)
    if sp.out_nchw:
        out = torch.zeros(sp.n, sp.cout, sp.hout, sp.wout, device=x.device)
        mode, out_c = ops.OUT_NCHW, 0
    else:
        out = ops.alloc_pnhwc(sp.n, sp.hout, sp.wout, sp.cpad_out, x.device, full=(sp.cout != sp.cpad_out), pooled=True)
        mode, out_c = ops.OUT_PNHWC, sp.cpad_out
    ops.conv2d_fused(cw, [ops.conv_src(x, sp.cpad_in, sp.taps)], sp.n, sp.hout, sp.wout, out, out_mode=mode, out_c=out_c,
                     hp=sp.hp, wp=sp.wp, bias=bias, addvec=(addvec[0] if addvec is not None else None),
                     addvec_stride=(addvec[1] if addvec is not None else 0), res=res, out_scale=sp.out_scale, splitk_ws=_splitk_ws(x.device))
    return out


def _embed(t, hp, wp):
    """Place a PNHWC tensor into a (possibly larger) zero padded space [N,hp,wp,C] (top-left aligned)."""
    if t.shape[1] == hp and t.shape[2] == wp:
        return t
    return F.pad(t, (0, 0, 0, wp - t.shape[2], 0, hp - t.shape[1]))


def _crop(t, hp, wp):
    return t if (t.shape[1] == hp and t.shape[2] == wp) else t[:, :hp, :wp, :].contiguous()


def _grad_target(p):
    """The .grad buffer a kernel may accumulate into directly, or None."""
    if ACCUM['on'] and not torch.is_grad_enabled() and isinstance(p, torch.nn.Parameter) and p.grad is not None \
            and p.grad.is_contiguous() and p.requires_grad:
        return p.grad
    return None


def _wgrad_into(dw, x, dye, sp):
    ops.conv_wgrad(x, dye, dw, sp.n, sp.hp, sp.wp, sp.cout, sp.cin, sp.cpad_in, sp.taps, sp.s_co, sp.s_ci, sp.s_tap,
                   precision=sp.prec, gain=sp.out_scale)


class ConvFn(Function):
    """y = sp.out_scale * (conv(x; w) + bias + addvec[n, c] + res)   (x, y, res PNHWC; y NCHW when sp.out_nchw).
    The scale and the residual ride in the conv epilogue (layerspp.py:307-309 "(x + h) / sqrt(2)"); in the backward the scale is
    the epilogue scale of the dgrad launch and the gain of the wgrad reduction, so it never costs an elementwise pass."""

    @staticmethod
    def forward(ctx, x, w, bias, addvec, res, sp):
        x = x.contiguous()
        ctx.w_ref, ctx.b_ref = w, bias
        w = w.contiguous()
        ctx.sp = sp
        ctx.has_bias, ctx.has_addvec, ctx.has_res = bias is not None, addvec is not None, res is not None
        ctx.save_for_backward(x, w)
        av = getattr(sp, 'av', None)      # (hub, column offset): addvec is the wide hub tensor, this conv reads / differentiates a slice
        if av is not None:
            assert addvec.stride(1) == 1
            addvec = (addvec.data_ptr() + 4 * av[1], addvec.stride(0))
        elif addvec is not None:
            addvec = addvec.contiguous()
            addvec = (addvec, addvec.shape[1])
        return _conv_forward(x, w, bias, addvec, sp, res.contiguous() if res is not None else None, w_ref=ctx.w_ref)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        sp = ctx.sp
        if sp.out_nchw:
            dy = ToPnhwcFn.apply(dy, sp.cpad_out)
        dyc = dy.contiguous()
        dx = DgradFn.apply(dyc, w, sp, ctx.w_ref) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            tgt = _grad_target(ctx.w_ref)
            if tgt is not None and not dyc.requires_grad:
                _wgrad_into(tgt, x, _embed(dyc, sp.hp, sp.wp).contiguous(), sp)
            else:
                dw = WgradFn.apply(x, dyc, sp)
        need_db = ctx.has_bias and ctx.needs_input_grad[2]
        need_dav = ctx.has_addvec and ctx.needs_input_grad[3]
        db = dav = dres = None
        if need_db or need_dav:
            if torch.is_grad_enabled() and dyc.requires_grad:
                # create_graph pass: keep the reductions differentiable
                assert getattr(sp, 'av', None) is None, 'hub-mode addvec is first-order only (generator)'
                db = dyc.sum(dim=(0, 1, 2))[:sp.cout] * sp.out_scale if need_db else None
                dav = dyc.sum(dim=(1, 2))[:, :sp.cout] * sp.out_scale if need_dav else None
            else:
                # one pass over dy gives the per-(sample, channel) sums (Dense_0 gradient) and the bias gradient, the latter
                # accumulated straight into bias.grad when that is allowed
                btgt = _grad_target(ctx.b_ref) if need_db else None
                av = getattr(sp, 'av', None)
                into = None
                if need_dav and av is not None:
                    into = (av[0].buf.data_ptr() + 4 * av[1], av[0].buf.stride(0))
                dav, db = ops.channel_grads(dyc, sp.cout, sp.out_scale, need_dav=need_dav, db_accum=btgt,
                                            need_db=need_db and btgt is None, dav_into=into)
        if ctx.has_res and ctx.needs_input_grad[4]:
            dres = dyc * sp.out_scale if sp.out_scale != 1.0 else dyc
        return dx, dw, db, dav, dres, None


class DgradFn(Function):
    """dx[q] = sum_t W_t^T dy[q - off_t]: the same implicit-GEMM kernel with adjoint taps and the transposed operand."""

    @staticmethod
    def forward(ctx, dy, w, sp, w_ref=None):
        ctx.sp = sp
        ctx.w_ref = w_ref
        dy = dy.contiguous()
        w = w.contiguous()
        ctx.save_for_backward(dy, w)
        # output space = the conv's input space (hp x wp padded); dy is embedded into it when the spaces differ (2x2-tap conv)
        dye = _embed(dy, sp.hp, sp.wp).contiguous()
        cy = dye.shape[-1]
        cw = _packed(w, w_ref, 'dgrad', sp, dy.device, cy=cy)
        dx = ops.alloc_pnhwc(sp.n, sp.hp - 2, sp.wp - 2, sp.cpad_in, dy.device, full=False, pooled=True)
        taps = [(-dr, -ds) for dr, ds in sp.taps]
        ops.conv2d_fused(cw, [ops.conv_src(dye, cy, taps)], sp.n, sp.hp - 2, sp.wp - 2, dx, out_scale=sp.out_scale,
                         splitk_ws=_splitk_ws(dy.device))
        if sp.cin < sp.cpad_in:
            dx[..., sp.cin:] = 0
        return dx

    @staticmethod
    def backward(ctx, ggx):
        dy, w = ctx.saved_tensors
        sp = ctx.sp
        ggx = ggx.contiguous()
        g_dy = g_w = None
        if ctx.needs_input_grad[0]:
            spf = SimpleNamespace(**{**vars(sp), 'out_nchw': False})
            g_dy = ConvFn.apply(ggx, ctx.w_ref if ctx.w_ref is not None else w, None, None, None, spf)
        if ctx.needs_input_grad[1]:
            tgt = _grad_target(ctx.w_ref)
            if tgt is not None and not ggx.requires_grad:
                _wgrad_into(tgt, ggx, _embed(dy, sp.hp, sp.wp).contiguous(), sp)
            else:
                g_w = WgradFn.apply(ggx, dy, sp)
        return g_dy, g_w, None, None


class WgradFn(Function):
    """dw[co][ci][t] = sum_q dy[q][co] x[q + off_t][ci]  (tcgen05 wgrad kernel, split-K with fp32 reductions)."""

    @staticmethod
    def forward(ctx, x, dy, sp):
        ctx.sp = sp
        x = x.contiguous()
        dy = dy.contiguous()
        ctx.save_for_backward(x, dy)
        dye = _embed(dy, sp.hp, sp.wp).contiguous()
        dw = torch.zeros(sp.w_shape, device=x.device)
        _wgrad_into(dw, x, dye, sp)
        return dw

    @staticmethod
    def backward(ctx, gdw):
        x, dy = ctx.saved_tensors
        sp = ctx.sp
        gdw = gdw.contiguous()
        g_x = g_dy = None
        if ctx.needs_input_grad[0]:
            g_x = DgradFn.apply(dy, gdw, sp, None)
        if ctx.needs_input_grad[1]:
            spf = SimpleNamespace(**{**vars(sp), 'out_nchw': False})
            g_dy = _crop(ConvFn.apply(x, gdw, None, None, None, spf), dy.shape[1], dy.shape[2])
        return g_x, g_dy, None


def _pad_cout(w, b, out_nchw=False):
    """PNHWC outputs carry pad_c(Cout) channels and the kernel writes whole 32-channel chunks: give odd Cout zero weight rows
    (differentiable: the gradient of F.pad slices them off again)."""
    cout = w.shape[0]
    cp = ops.pad_c(cout)
    if out_nchw or cp == cout:
        return w, b
    w = F.pad(w, (0, 0) * (w.dim() - 1) + (0, cp - cout))
    if b is not None:
        b = F.pad(b, (0, cp - cout))
    return w, b


def conv3x3(x, w, b, n, h, wd, addvec=None, out_nchw=False, prec=3, res=None, out_scale=1.0, av=None):
    """nn.Conv2d(k=3, s=1, p=1) on PNHWC; w [Cout, Cin, 3, 3] (zero-padded along Cin to x's channel count if needed).
    addvec: [N, Cout] per-sample add, or -- with av = (hub, column offset) -- the wide GradHub tensor holding it as a slice."""
    cout_real = w.shape[0]
    w, b = _pad_cout(w, b, out_nchw)
    if addvec is not None and w.shape[0] != cout_real:
        assert av is None
        addvec = F.pad(addvec, (0, w.shape[0] - cout_real))
    cout, cin = w.shape[0], w.shape[1]
    cp = x.shape[-1]
    if cin < cp:
        w = F.pad(w, (0, 0, 0, 0, 0, cp - cin))
    sp = conv_spec(w.shape, ops.TAPS_3X3, n, h, wd, cp, cp * 9, 9, 1, cout, cp, out_nchw=out_nchw, prec=prec, out_scale=out_scale)
    sp.av = av
    return ConvFn.apply(x, w, b, addvec, res, sp)


def conv1x1(x, w, b, n, h, wd, prec=3, res=None, out_scale=1.0):
    """1x1 conv; w [Cout, Cin, 1, 1] or [Cout, Cin]."""
    w, b = _pad_cout(w, b)
    cout, cin = w.shape[0], w.shape[1]
    cp = x.shape[-1]
    if cin < cp:
        w = F.pad(w.reshape(cout, cin), (0, cp - cin))
    # a [Cout, Cin, 1, 1] leaf is passed as it is (same memory as [Cout, Cin]): its .grad can then be accumulated in place
    sp = conv_spec(w.shape, ops.TAPS_1X1, n, h, wd, cp, cp, 1, 0, cout, cp, prec=prec, out_scale=out_scale)
    return ConvFn.apply(x, w, b, None, res, sp)


def nin(x, W, b, n, h, wd, prec=3, res=None, out_scale=1.0):
    """layers.py:489-512: weight stored [in, out]."""
    cin, cout = W.shape
    sp = conv_spec(W.shape, ops.TAPS_1X1, n, h, wd, x.shape[-1], 1, cout, 0, cout, cin, prec=prec, out_scale=out_scale)
    return ConvFn.apply(x, W, b, None, res, sp)


# ------------------------------------------------------------------------------------------------------------------
# FIR resampling
# ------------------------------------------------------------------------------------------------------------------
class FirFn(Function):
    """mode 1: up x2 (gain 4 built in), 2: down x2, 3: pad(2,2) + space-to-depth, 4: adjoint of 3; `gain` scales the taps."""

    @staticmethod
    def forward(ctx, x, mode, gain, c_img):
        ctx.mode, ctx.gain, ctx.c_img = mode, gain, c_img
        x = x.contiguous()
        n, hp, wp, c = x.shape
        if mode == 1:
            out = ops.alloc_pnhwc(n, 2 * (hp - 2), 2 * (wp - 2), c, x.device, full=False, pooled=True)
        elif mode == 2:
            out = ops.alloc_pnhwc(n, (hp - 2) // 2, (wp - 2) // 2, c, x.device, full=False, pooled=True)
        elif mode == 3:
            out = torch.zeros(n, (hp - 2) // 2 + 3, (wp - 2) // 2 + 3, 4 * c, device=x.device)
        else:
            out = ops.alloc_pnhwc(n, 2 * (hp - 3), 2 * (wp - 3), c_img, x.device)
        ctx.in_c = c
        ops.fir_pnhwc(x, mode, out, gain=gain)
        return out

    @staticmethod
    def backward(ctx, g):
        adj = {1: (2, 4.0 * ctx.gain), 2: (1, 0.25 * ctx.gain), 3: (4, ctx.gain), 4: (3, ctx.gain)}[ctx.mode]
        return FirFn.apply(g, adj[0], adj[1], ctx.in_c), None, None, None


def fir_up(x):
    return FirFn.apply(x, 1, 1.0, 0)


def fir_down(x):
    return FirFn.apply(x, 2, 1.0, 0)


# ------------------------------------------------------------------------------------------------------------------
# GroupNorm = stats + [N,C] algebra + affine/activation
# ------------------------------------------------------------------------------------------------------------------
class StatsFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.stats_fwd(x)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return ops.stats_bwd(x, g.to(torch.float32).contiguous())


class AffineActFn(Function):
    @staticmethod
    def forward(ctx, x, scale, shift, act):
        x, scale, shift = x.contiguous(), scale.contiguous(), shift.contiguous()
        ctx.act = act
        ctx.save_for_backward(x, scale, shift)
        return ops.affine_act_fwd(x, scale, shift, act)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, scale, shift = ctx.saved_tensors
        dx, sums = ops.affine_act_bwd(x, dy.contiguous(), scale, shift, ctx.act)
        return dx, sums[..., 0].to(torch.float32), sums[..., 1].to(torch.float32), None


class GnCoefFn(Function):
    """(scale, shift)[N, C] of a GroupNorm from the per-(sample, channel) sums: one kernel each way (ddg_gn_prepare / _bwd)
    instead of ~45 small float64 torch kernels per normalisation.  First-order only: the generator's norms never sit under
    the R1 double backward (the discriminator has none)."""

    @staticmethod
    def forward(ctx, st, gamma, beta, groups, hw, eps):
        n, c = st.shape[0], st.shape[1]
        per_sample = gamma.dim() == 2
        assert gamma.stride(-1) == 1 and beta.stride(-1) == 1
        gb_stride = gamma.stride(0) if per_sample else 0
        if per_sample:
            assert beta.stride(0) == gb_stride
        st = st.contiguous()
        scale = torch.empty(n, c, device=st.device, dtype=torch.float32)
        shift = torch.empty(n, c, device=st.device, dtype=torch.float32)
        ops.gn_prepare(st, c, None, 0, gamma, beta, gb_stride, per_sample, n, hw, groups, scale, shift, eps)
        ctx.save_for_backward(st, gamma)
        ctx.cfg = (groups, hw, eps, per_sample, gb_stride)
        return scale, shift

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dscale, dshift):
        st, gamma = ctx.saved_tensors
        groups, hw, eps, per_sample, gb_stride = ctx.cfg
        n, c = st.shape[0], st.shape[1]
        dst, dga, dbe = ops.gn_prepare_bwd(st, gamma, gb_stride, per_sample, dscale.contiguous(), dshift.contiguous(), n, c, hw,
                                           groups, eps)
        if not per_sample:
            dga, dbe = dga.sum(0), dbe.sum(0)
        return dst, dga, dbe, None, None, None


def group_norm_act(x, h, w, groups, gamma, beta, act, eps=1e-6):
    """gamma / beta: [N, C] (AdaGN) or [C] (affine GroupNorm).  Statistics in float64 as in the fused plan."""
    scale, shift = GnCoefFn.apply(StatsFn.apply(x), gamma, beta, groups, h * w, eps)
    return AffineActFn.apply(x, scale, shift, act)


class GradHub(Function):
    """Identity on a wide [N, J] tensor whose consumers each read one column slice (the batched AdaGN style projections and the
    batched Dense_0 projections).  Instead of J/slice SliceBackward nodes (a zero-filled [N, J] temporary, a copy and an add
    each: ~250 launches per step) the consumers write their slice of the gradient straight into `hub.buf` in their own backward
    and hand autograd None; this node returns the assembled buffer once all of them have run."""

    @staticmethod
    def forward(ctx, x, hub):
        ctx.set_materialize_grads(False)
        hub.buf = torch.zeros_like(x)
        ctx.hub = hub
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        buf = ctx.hub.buf
        return (buf if g is None else buf + g), None


class _Arena:
    """Zeroed float64 scratch for the per-(sample, channel) statistics of one forward (+ the first-pass sums of its backward):
    one fill per forward instead of two per GroupNorm.  Capacity is learned from the previous forward of the module."""

    def __init__(self, cap, device):
        self.buf = torch.zeros(cap, dtype=torch.float64, device=device) if cap > 0 else None
        self.cap, self.used, self.want = cap, 0, 0

    def take(self, n, device):
        self.want += n
        if self.used + n <= self.cap:
            v = self.buf[self.used:self.used + n]
            self.used += n
            return v
        return torch.zeros(n, dtype=torch.float64, device=device)


_CUR_ARENA = [None]


def _begin_arena(mod, device):
    ar = _Arena(getattr(mod, '_arena_cap', 0), device)
    _CUR_ARENA[0] = ar
    return ar


def _end_arena(mod, ar):
    mod._arena_cap = max(getattr(mod, '_arena_cap', 0), 2 * ar.want)    # forward statistics + as many backward sums
    _CUR_ARENA[0] = None


class GnActFn(Function):
    """y = act(gamma[n,c] * GroupNorm(x) + beta[n,c]) on PNHWC as one autograd node (layerspp.py:46-63 + the SiLU of :279,300).
    forward : per-(n,c) sums -> scale/shift (ddg_gn_prepare) -> apply;  3 launches
    backward: first-pass sums -> d(gamma), d(beta), statistics coefficients (ddg_gn_bwd_coeffs) -> dx in one pass;  3 launches
    (was: 3 + 5 kernels, three full-size temporaries and their border clears, fp64 -> fp32 casts).
    g: SimpleNamespace(groups, hw, eps, act, per_sample, off, c, hub, arena): with per_sample the affine parameters are columns
    [off, off+c) / [off+c, off+2c) of the wide style tensor `gsrc` and their gradient goes to hub.buf; else gsrc / bsrc are the
    [C] affine parameters of a plain GroupNorm."""

    @staticmethod
    def forward(ctx, x, gsrc, bsrc, g):
        x = x.contiguous()
        n, c = x.shape[0], x.shape[-1]
        dev = x.device
        ar = g.arena
        st = (ar.take(n * c * 2, dev) if ar is not None else torch.zeros(n * c * 2, dtype=torch.float64, device=dev))
        ops.stats_fwd(x, out=st)
        if g.per_sample:
            assert gsrc.stride(1) == 1
            gb_stride = gsrc.stride(0)
            gptr, bptr = gsrc.data_ptr() + 4 * g.off, gsrc.data_ptr() + 4 * (g.off + c)
        else:
            gb_stride = 0
            gptr, bptr = gsrc.contiguous().data_ptr(), bsrc.contiguous().data_ptr()
        scale = torch.empty(n, c, device=dev, dtype=torch.float32)
        shift = torch.empty(n, c, device=dev, dtype=torch.float32)
        ops.gn_prepare(st, c, None, 0, gptr, bptr, gb_stride, g.per_sample, n, g.hw, g.groups, scale, shift, g.eps)
        ctx.g = g
        ctx.gb = (gb_stride, g.off)
        ctx.save_for_backward(x, scale, shift, st, gsrc, bsrc)
        return ops.affine_act_fwd(x, scale, shift, g.act)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, scale, shift, st, gsrc, bsrc = ctx.saved_tensors
        g = ctx.g
        n, c = x.shape[0], x.shape[-1]
        dev = x.device
        dy = dy.contiguous()
        ar = g.arena
        fresh = ar is not None and not getattr(ctx, 'ran', False)
        sums = ar.take(n * c * 2, dev) if fresh else torch.zeros(n * c * 2, dtype=torch.float64, device=dev)
        ctx.ran = True
        ops.affine_act_bwd(x, dy, scale, shift, g.act, need_dx=False, sums=sums)
        g12 = torch.empty(n, c, 2, device=dev, dtype=torch.float32)
        gb_stride, off = ctx.gb
        if g.per_sample:
            gptr = gsrc.data_ptr() + 4 * off
            if g.hub is not None:
                base = g.hub.buf.data_ptr()
                ops.gn_bwd_coeffs(st, sums, gptr, gb_stride, True, g12, base + 4 * off, base + 4 * (off + c), g.hub.buf.stride(0), n, c,
                                  g.hw, g.groups, g.eps)
                d_g = d_b = None
            else:
                d_g = torch.zeros_like(gsrc)
                ops.gn_bwd_coeffs(st, sums, gptr, gb_stride, True, g12, d_g.data_ptr() + 4 * off, d_g.data_ptr() + 4 * (off + c),
                                  d_g.stride(0), n, c, g.hw, g.groups, g.eps)
                d_b = None
        else:
            dga = torch.empty(n, c, device=dev, dtype=torch.float32)
            dbe = torch.empty(n, c, device=dev, dtype=torch.float32)
            ops.gn_bwd_coeffs(st, sums, gsrc.contiguous(), 0, False, g12, dga, dbe, c, n, c, g.hw, g.groups, g.eps)
            d_g, d_b = dga.sum(0), dbe.sum(0)
        dx = ops.gn_bwd_dx(x, dy, scale, shift, g12, g.act)
        return dx, d_g, d_b, None


def gn_act(x, h, w, groups, act, gamma=None, beta=None, style=None, off=0, hub=None, eps=1e-6):
    """Fused GroupNorm (+ AdaGN affine) + activation.  Either gamma / beta ([C] parameters) or style = the wide [N, J] projection
    tensor with this norm's [gamma | beta] at columns [off, off + 2C)."""
    g = SimpleNamespace(groups=groups, hw=h * w, eps=eps, act=act, per_sample=style is not None, off=off, c=x.shape[-1], hub=hub,
                        arena=_CUR_ARENA[0])
    if style is not None:
        return GnActFn.apply(x, style, None, g)
    return GnActFn.apply(x, gamma, beta, g)


def up_first_idx(cfg):
    """all_modules index of the first up-path block (same walk as generator_forward): the modules from there on are the ones
    differentiated first in the backward pass."""
    cfg = arch.normalize_config(cfg)
    rest = arch.ncsnpp_modules(cfg)[3:]
    i, cur, nres = 0, cfg.image_size, len(cfg.ch_mult)
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            i += 1
            if cur in cfg.attn_resolutions:
                i += 1
        if lvl != nres - 1:
            i += 1
            if cfg.progressive_input == 'residual':
                i += 1
            cur //= 2
    return rest[i + 3]['idx']


class AttnCoreFn(Function):
    """o = softmax(q k^T / sqrt(C)) v per sample (layerspp.py:115-119) on this library's kernels, forward and backward: batched
    tcgen05 GEMMs with per-image packed operands (ops.bgemm), ddg_softmax_rows and its backward.  q, k, v: [N, T, C] fp32.
    First-order only (the generator never sits under the R1 double backward)."""

    @staticmethod
    def forward(ctx, q, k, v, prec):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        n, t, c = q.shape
        tp = ops.pad_c(t)
        scale = float(c) ** -0.5
        s = ops.bgemm(q, k, t, c, c, 1, t * c, precision=prec, out_scale=scale, out_c=tp)          # [N, T, Tp] logits
        p = torch.empty_like(s)
        ops.softmax_rows(s, p, n * t, t, tp, tp)
        o = ops.bgemm(p, v, c, t, 1, c, t * c, precision=prec)                                       # B[c][s] = v[s][c]
        ctx.save_for_backward(q, k, v, p)
        ctx.cfg = (prec, scale, tp)
        return o

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, do):
        q, k, v, p = ctx.saved_tensors
        prec, scale, tp = ctx.cfg
        n, t, c = q.shape
        do = do.contiguous()

        def transposed(m):            # [N, T, Tp] (valid [:, :, :T]) -> its per-image transpose, same padding
            if tp == t:
                return m.transpose(1, 2).contiguous()
            out = torch.zeros_like(m)
            out[:, :, :t] = m[:, :, :t].transpose(1, 2)
            return out
        dv = ops.bgemm(transposed(p), do, c, t, 1, c, t * c, precision=prec)                        # dV[s][c] = sum_t P[t][s] dO[t][c]
        dp = ops.bgemm(do, v, t, c, c, 1, t * c, precision=prec, out_c=tp)                           # dP[t][s] = sum_c dO[t][c] V[s][c]
        ds = torch.empty_like(dp)
        ops.softmax_rows_bwd(p, dp, ds, n * t, t, tp, scale)                                         # includes the 1/sqrt(C)
        dq = ops.bgemm(ds, k, c, t, 1, c, t * c, precision=prec)                                     # dQ[t][c] = sum_s dS[t][s] K[s][c]
        dk = ops.bgemm(transposed(ds), q, c, t, 1, c, t * c, precision=prec)                         # dK[s][c] = sum_t dS[t][s] Q[t][c]
        return dq, dk, dv, None


def _tc_linear(n, k, j):
    """Large projections (the batched AdaGN style / Dense_0 GEMMs) go to the tcgen05 kernels; small ones stay on the SIMT kernel."""
    return j >= 1024 and k % 32 == 0 and n % 8 == 0 and j % 32 == 0


class LinearFn(Function):
    """y = x W^T + b on [N, K] rows (forward and both gradient GEMMs): ddg_linear, or the tensor-core GEMMs for wide layers.
    Closed under differentiation: under create_graph the backward is expressed with LinearFn itself (dx = dy W is a linear
    layer with weight W^T, dW = dy^T x one with weight x^T), which the R1 penalty needs through the discriminator's
    end_linear (discriminator.py:165-167)."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        ctx.has_b = b is not None
        xc, Wc = x.contiguous(), W.contiguous()
        if _tc_linear(xc.shape[0], xc.shape[1], Wc.shape[0]):
            return ops.gemm_rows(xc, Wc, b)
        return ops.linear(xc, Wc, b)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dx = dW = None
        if torch.is_grad_enabled():
            # differentiable backward (double backward of the gradient penalty)
            if ctx.needs_input_grad[0]:
                dx = LinearFn.apply(dy, W.t(), None)
            if ctx.needs_input_grad[1]:
                dW = LinearFn.apply(dy.t(), x.t(), None)
            db = dy.sum(0) if (ctx.has_b and ctx.needs_input_grad[2]) else None
            return dx, dW, db
        dy = dy.contiguous()
        xc, Wc = x.contiguous(), W.contiguous()
        tc = _tc_linear(xc.shape[0], xc.shape[1], Wc.shape[0])
        if ctx.needs_input_grad[0]:
            # dx[n][k] = sum_j dy[n][j] W[j][k]: contraction over the J rows of dy^T and W
            dx = ops.gemm_tn(dy.t().contiguous(), Wc) if tc else ops.linear(dy, Wc.t().contiguous())
        if ctx.needs_input_grad[1]:
            # dW[j][k] = sum_n dy[n][j] x[n][k]: contraction over the batch rows
            dW = ops.gemm_tn(dy, xc) if tc else ops.linear(dy.t().contiguous(), xc.t().contiguous())
        db = dy.sum(0) if (ctx.has_b and ctx.needs_input_grad[2]) else None
        return dx, dW, db


def _groups(c):
    return min(c // 4, 32)


def _interior(x):
    return x[:, 1:-1, 1:-1, :]


class S2dWeightFn(Function):
    """wt [Cout, Cin, 3, 3] -> w2[co, py, px, ci, dy, dx] = wt[co, ci, 2dy+py, 2dx+px] (zero elsewhere); linear, so the backward
    is the adjoint gather and the double backward the forward again."""

    @staticmethod
    def forward(ctx, wt, cp):
        ctx.dims = (wt.shape[0], wt.shape[1], cp)
        return ops.s2d_weights(wt, wt.shape[0], wt.shape[1], cp)

    @staticmethod
    def backward(ctx, g):
        cout, cin, cp = ctx.dims
        return S2dWeightAdjFn.apply(g, cout, cin, cp), None


class S2dWeightAdjFn(Function):
    @staticmethod
    def forward(ctx, g, cout, cin, cp):
        ctx.cp = cp
        return ops.s2d_weights(g, cout, cin, cp, adjoint=True)

    @staticmethod
    def backward(ctx, gg):
        return S2dWeightFn.apply(gg, ctx.cp), None, None, None


def conv_downsample_pnhwc(x, wt, bias, n, h, w, prec=3, res=None, out_scale=1.0):
    """up_or_down_sampling.py:226-262 conv_downsample_2d with k = [1,3,3,1], factor 2 on a PNHWC tensor: pad(2,2) FIR written
    space-to-depth (FirFn mode 3), then the stride-2 3x3 conv as a stride-1 2x2-tap conv over the 4*C s2d channels.
    x [N, h+2, w+2, cp]; wt [Cout, Cin, 3, 3]; returns PNHWC [N, h/2+2, w/2+2, pad_c(Cout)] = out_scale * (conv + bias + res)."""
    cp = x.shape[-1]
    ho, wo = h // 2, w // 2
    s2d = FirFn.apply(x, 3, 1.0, 0)                           # [N, ho+3, wo+3, 4*cp]
    wt, bias = _pad_cout(wt, bias)
    cout, cin = wt.shape[0], wt.shape[1]
    w2 = S2dWeightFn.apply(wt, cp)                            # [Cout, 2, 2, cp, 2, 2]
    w2 = w2.reshape(cout, 4 * cp, 4)
    sp = conv_spec(w2.shape, ops.TAPS_2X2, n, ho, wo, 4 * cp, 4 * cp * 4, 4, 1, cout, 4 * cp, hp=ho + 3, wp=wo + 3, prec=prec,
                   out_scale=out_scale)
    return ConvFn.apply(s2d, w2, bias, None, res, sp)


# ------------------------------------------------------------------------------------------------------------------
# generator
# ------------------------------------------------------------------------------------------------------------------
def generator_forward(mod, x, time_cond, z):
    """NCSNpp.forward (ncsnpp_generator_adagn.py:280-431), training path.  `mod` is ddgan_b200.modules.NCSNpp."""
    cfg = mod.cfg
    if not (cfg.resblock_type == 'biggan' and cfg.embedding_type == 'positional' and cfg.progressive == 'none'
            and cfg.progressive_input in ('residual', 'none') and cfg.fir and cfg.conditional):
        raise NotImplementedError('training path covers the biggan / positional / fir configuration family')
    if cfg.num_channels_dae % 32 != 0:
        raise NotImplementedError('training path needs num_channels_dae % 32 == 0 (channel-concatenated skips are not padded)')
    P = dict(mod.named_parameters())
    prec = mod.precision
    N, S = x.shape[0], cfg.image_size
    nf = cfg.num_channels_dae
    packs = train_packs(mod)
    packs.begin()
    _CUR_PACKS[0] = packs
    drop = float(cfg.dropout) if mod.training else 0.0

    # z mapping and time embedding
    zn = z / torch.sqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    zemb = F.silu(LinearFn.apply(zn, P['z_transform.1.weight'], P['z_transform.1.bias']))
    for i in range(cfg.n_mlp):
        zemb = F.silu(LinearFn.apply(zemb, P[f'z_transform.{3 + 2 * i}.weight'], P[f'z_transform.{3 + 2 * i}.bias']))
    temb = ops.timestep_embedding(time_cond, nf)
    temb = LinearFn.apply(temb, P['all_modules.0.weight'], P['all_modules.0.bias'])
    temb = LinearFn.apply(F.silu(temb), P['all_modules.1.weight'], P['all_modules.1.bias'])
    temb_act = F.silu(temb)

    # all AdaGN style projections (layerspp.py:57) and all Dense_0 projections (layerspp.py:298-299) as two batched GEMMs
    style_names = [k[:-7] for k in P if k.endswith('.style.weight')]
    dense_names = [k[:-7] for k in P if k.endswith('Dense_0.weight')]
    style_all = LinearFn.apply(zemb, torch.cat([P[k + '.weight'] for k in style_names], 0),
                               torch.cat([P[k + '.bias'] for k in style_names], 0))
    dense_all = LinearFn.apply(temb_act, torch.cat([P[k + '.weight'] for k in dense_names], 0),
                               torch.cat([P[k + '.bias'] for k in dense_names], 0))
    style_off, dense_off = {}, {}
    o = 0
    for k in style_names:
        style_off[k] = o; o += P[k + '.weight'].shape[0]
    o = 0
    for k in dense_names:
        dense_off[k] = o; o += P[k + '.weight'].shape[0]
    style_hub, dense_hub = SimpleNamespace(buf=None), SimpleNamespace(buf=None)
    style_all = GradHub.apply(style_all, style_hub)
    dense_all = GradHub.apply(dense_all, dense_hub)
    arena = _begin_arena(mod, x.device)

    def adagn(t, h, w, prefix, act=ops.ACT_SILU):
        c = t.shape[-1]
        return gn_act(t, h, w, _groups(c), act, style=style_all, off=style_off[prefix + '.style'], hub=style_hub)

    def resblock(m, t, h, w):
        pn = f"all_modules.{m['idx']}."
        cin, cout = m['i'], m['o']
        hh = adagn(t, h, w, pn + 'GroupNorm_0')
        xs = t
        if m['up']:
            hh, xs, h, w = fir_up(hh), fir_up(xs), 2 * h, 2 * w
        elif m['down']:
            hh, xs, h, w = fir_down(hh), fir_down(xs), h // 2, w // 2
        hh = conv3x3(hh, P[pn + 'Conv_0.weight'], P[pn + 'Conv_0.bias'], N, h, w, addvec=dense_all, prec=prec,
                     av=(dense_hub, dense_off[pn + 'Dense_0']))
        hh = adagn(hh, h, w, pn + 'GroupNorm_1')
        if drop > 0:
            hh = F.dropout(hh, drop, True)
        if cin != cout or m['up'] or m['down']:
            xs = conv1x1(xs, P[pn + 'Conv_2.weight'], P[pn + 'Conv_2.bias'], N, h, w, prec=prec)
        # (skip + h) / sqrt(2) in Conv_1's epilogue
        out = conv3x3(hh, P[pn + 'Conv_1.weight'], P[pn + 'Conv_1.bias'], N, h, w, prec=prec, res=xs,
                      out_scale=RSQRT2 if cfg.skip_rescale else 1.0)
        return out, h, w

    def attn(m, t, h, w):
        pn = f"all_modules.{m['idx']}."
        c = t.shape[-1]
        g = gn_act(t, h, w, _groups(c), ops.ACT_NONE, gamma=P[pn + 'GroupNorm_0.weight'], beta=P[pn + 'GroupNorm_0.bias'])
        q = _interior(nin(g, P[pn + 'NIN_0.W'], P[pn + 'NIN_0.b'], N, h, w, prec)).reshape(N, h * w, c)
        k = _interior(nin(g, P[pn + 'NIN_1.W'], P[pn + 'NIN_1.b'], N, h, w, prec)).reshape(N, h * w, c)
        v = _interior(nin(g, P[pn + 'NIN_2.W'], P[pn + 'NIN_2.b'], N, h, w, prec)).reshape(N, h * w, c)
        o = AttnCoreFn.apply(q, k, v, prec).reshape(N, h, w, c)
        o = F.pad(o, (0, 0, 1, 1, 1, 1))
        return nin(o, P[pn + 'NIN_3.W'], P[pn + 'NIN_3.b'], N, h, w, prec, res=t, out_scale=RSQRT2 if cfg.skip_rescale else 1.0)

    def pyramid_down(m, pyr, hcur, h, w):
        """conv_downsample_2d + bias, then (pyramid + h)/sqrt2 (ncsnpp...:343-350); h, w = input size of pyr."""
        pn = f"all_modules.{m['idx']}.Conv2d_0"
        return conv_downsample_pnhwc(pyr, P[pn + '.weight'], P[pn + '.bias'], N, h, w, prec=prec, res=hcur,
                                     out_scale=RSQRT2 if cfg.skip_rescale else 1.0)

    mods = arch.ncsnpp_modules(cfg)
    xin = x if cfg.centered else 2 * x - 1.0
    xp = ToPnhwcFn.apply(xin, ops.pad_c(cfg.num_channels))
    m0 = mods[2]
    h = conv3x3(xp, P[f"all_modules.{m0['idx']}.weight"], P[f"all_modules.{m0['idx']}.bias"], N, S, S, prec=prec)
    hs = [(h, S)]
    pyramid, pyr_size = xp, S
    rest = mods[3:]
    i = 0
    nres = len(cfg.ch_mult)
    cur = S
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            m = rest[i]; i += 1
            h, _, _ = resblock(m, hs[-1][0], cur, cur)
            if cur in cfg.attn_resolutions:
                h = attn(rest[i], h, cur, cur); i += 1
            hs.append((h, cur))
        if lvl != nres - 1:
            m = rest[i]; i += 1
            h, nh, _ = resblock(m, hs[-1][0], cur, cur)
            if cfg.progressive_input == 'residual':
                m = rest[i]; i += 1
                h = pyramid_down(m, pyramid, h, pyr_size, pyr_size)
                pyramid, pyr_size = h, nh
            cur = nh
            hs.append((h, cur))
    h = hs[-1][0]
    h, _, _ = resblock(rest[i], h, cur, cur); i += 1
    h = attn(rest[i], h, cur, cur); i += 1
    h, _, _ = resblock(rest[i], h, cur, cur); i += 1
    # Data-parallel overlap point: when the backward pass gets back here, every weight gradient of the up path and the head (module
    # index >= up_first_idx(cfg), about 60 % of the parameters) has been enqueued -- train.Trainer all-reduces that part of the
    # gradient arena on a side stream while the down path is still being differentiated.
    hook = GRAD_READY_HOOKS.get(mod)
    hook = hook() if hook is not None else None          # WeakMethod: the Trainer may be gone
    if hook is not None and torch.is_grad_enabled() and h.requires_grad:
        h.register_hook(lambda g, hook=hook: hook())
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            skip, _ = hs.pop()
            h, _, _ = resblock(rest[i], torch.cat([h, skip], dim=-1), cur, cur); i += 1
        if cur in cfg.attn_resolutions:
            h = attn(rest[i], h, cur, cur); i += 1
        if lvl != 0:
            h, cur, _ = resblock(rest[i], h, cur, cur); i += 1
    assert not hs
    g, c = rest[i], rest[i + 1]
    h = gn_act(h, cur, cur, _groups(h.shape[-1]), ops.ACT_SILU, gamma=P[f"all_modules.{g['idx']}.weight"],
               beta=P[f"all_modules.{g['idx']}.bias"])
    _end_arena(mod, arena)
    y = conv3x3(h, P[f"all_modules.{c['idx']}.weight"], P[f"all_modules.{c['idx']}.bias"], N, cur, cur, out_nchw=True, prec=prec)
    _CUR_PACKS[0] = None
    return y if cfg.not_use_tanh else torch.tanh(y)


# ------------------------------------------------------------------------------------------------------------------
# discriminators
# ------------------------------------------------------------------------------------------------------------------
def discriminator_forward(mod, x, t, x_t):
    """Discriminator_small/large.forward (discriminator.py:134-167 / :205-238), training path (double-differentiable)."""
    if (8 * mod.ngf) % 32 != 0:
        raise NotImplementedError('training path needs ngf % 4 == 0 (the stddev channel is appended after 8*ngf unpadded channels)')
    P = dict(mod.named_parameters())
    prec = mod.precision
    N, S = x.shape[0], x.shape[-1]
    packs = train_packs(mod)
    packs.begin()
    _CUR_PACKS[0] = packs
    te = ops.timestep_embedding(t, mod.t_emb_dim)
    te = LinearFn.apply(te, P['t_embed.main.0.weight'], P['t_embed.main.0.bias'])
    te = LinearFn.apply(F.leaky_relu(te, 0.2), P['t_embed.main.2.weight'], P['t_embed.main.2.bias'])
    te = F.leaky_relu(te, 0.2)
    xin = ToPnhwcFn.apply(torch.cat((x, x_t), dim=1), ops.pad_c(mod.nc))
    h = conv1x1(xin, P['start_conv.weight'], P['start_conv.bias'], N, S, S, prec=prec)
    cur = S
    for i, (a, b, ds) in enumerate(arch.discriminator_blocks(mod.ngf, mod.large)):
        pn = f'conv{i + 1}.'
        dense = LinearFn.apply(te, P[pn + 'dense_t1.weight'], P[pn + 'dense_t1.bias'])
        o = conv3x3(F.leaky_relu(h, 0.2), P[pn + 'conv1.0.weight'], P[pn + 'conv1.0.bias'], N, cur, cur, addvec=dense, prec=prec)
        o = F.leaky_relu(o, 0.2)
        xs = h
        if ds:
            o, xs, cur = fir_down(o), fir_down(xs), cur // 2
        sk = conv1x1(xs, P[pn + 'skip.0.weight'], None, N, cur, cur, prec=prec)
        h = conv3x3(o, P[pn + 'conv2.0.weight'], P[pn + 'conv2.0.bias'], N, cur, cur, prec=prec, res=sk, out_scale=RSQRT2)
    # minibatch stddev (discriminator.py:150-158) on the interior, appended as one extra (32-padded) channel group
    c = h.shape[-1]
    group = min(N, mod.stddev_group)
    inner = _interior(h)                                            # [N, H, W, C]
    sd = inner.reshape(group, -1, cur, cur, c)
    sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8).mean(dim=(1, 2, 3))   # [N/group]
    sd = sd.repeat(group)                                           # sample i -> stat-set i % (N/group)
    extra = h.new_zeros(N, cur, cur, 32)
    extra = torch.cat([sd.view(N, 1, 1, 1).expand(N, cur, cur, 1), extra[..., 1:]], dim=-1)
    extra = F.pad(extra, (0, 0, 1, 1, 1, 1))
    hcat = torch.cat([h, extra], dim=-1)
    f = conv3x3(hcat, P['final_conv.weight'], P['final_conv.bias'], N, cur, cur, prec=prec)
    f = F.leaky_relu(_interior(f), 0.2).sum(dim=(1, 2))            # [N, C]
    _CUR_PACKS[0] = None
    return LinearFn.apply(f, P['end_linear.weight'], P['end_linear.bias'])
