"""Static launch plans that lower the reference's model forwards onto the fused sm_100a kernels.

`GeneratorEngine` is NCSNpp.forward (score_sde/models/ncsnpp_generator_adagn.py:280-431) for the configuration
family the reference's README trains (resblock_type='biggan', embedding_type='positional', progressive='none',
progressive_input in {'residual','none'}, fir=True); `DiscriminatorEngine` is Discriminator_small/large.forward
(score_sde/models/discriminator.py:134-167, 205-238).

Design (B200-first, not a module-by-module translation):
  * activations live in padded NHWC ("PNHWC", zero border) so a 3x3 tap is a pointer offset for the tcgen05 kernel;
  * every GroupNorm/AdaGN + SiLU is folded into the *consumer* conv's prologue (scale/shift per (n,c)); its statistics
    are accumulated by the *producer* conv's epilogue -> no normalisation kernel ever touches the activations;
  * bias, "+Dense_0(temb)", the 1x1 skip conv (extra K segments of the same accumulator), the residual add and the
    1/sqrt(2) rescale are epilogue / K-loop work of the second conv of a block;
  * all AdaGN style projections and all Dense_0 projections are two batched linear launches per forward;
  * the attention GEMMs run on the same tcgen05 kernel in batched mode with per-image packed K / V operands;
  * parameters and activations are allocated once at build time (fixed addresses), so a forward -- and the T-step
    sampling loop around it -- is one CUDA graph that survives weight updates.
"""
from __future__ import annotations

import math
import os

import torch

from . import arch, ops

RSQRT2 = 1.0 / math.sqrt(2.0)


class Act:
    """One activation of the plan: PNHWC buffer + optional per-(n,c) {sum, sumsq} accumulators (float64)."""

    def __init__(self, eng, c, h, w, stats=True):
        self.C, self.H, self.W = c, h, w
        self.buf = ops.alloc_pnhwc(eng.N, h, w, c, eng.dev)
        self.stats = eng._alloc_stats(c) if stats else None
        self.planes = None      # pre-split bf16 copy (ops.alloc_planes), created on demand by _EngineBase._planes_of
        self.producer = None    # conv descriptor that writes this activation (PNHWC), if any: it can write the planes too


def _groups(c):
    return min(c // 4, 32)  # layerspp.py:254,267 / :100


class _EngineBase:
    def __init__(self, batch, device, precision, shapes, params=None):
        """params: optional {name: tensor} whose storage the engine reads directly (the drop-in modules pass their own
        parameters: no per-update copies; the caller rebuilds the engine if those tensors move).  Default: engine-owned
        fixed-address copies filled by load_state_dict."""
        self.N = batch
        self.dev = torch.device(device)
        self.prec = precision
        self.shapes = shapes
        if params is not None:
            self.P = {}
            for k, v in shapes.items():
                t = params[k].detach()
                assert tuple(t.shape) == tuple(v) and t.is_contiguous() and t.device == self.dev and t.dtype == torch.float32, k
                self.P[k] = t
        else:
            self.P = {k: torch.zeros(v, device=self.dev) for k, v in shapes.items()}  # fixed-address parameter copies
        self._pack_plan = None
        self._splitk_ws = ops.alloc_splitk_ws(self.dev)   # split-K workspace of the 4x4 / 8x8 level convs (one stream: shared by all)
        self.steps = []      # zero-arg callables, executed in order on the current stream
        self.step_names = []
        self.binders = []    # zero-arg callables that (re)pack derived weights from self.P
        self._stats_cap = 64 * 1024 * 1024 // 8
        self._stats_arena = torch.zeros(self._stats_cap, dtype=torch.float64, device=self.dev)
        self._stats_used = 0
        self._keep = []
        self.graph = None
        self.n_launches = 0
        self.conv_flops = 0
        self.conv_bytes = 0      # algorithmic HBM bytes of the conv launches: inputs + packed weights + output (+ residual), once each

    def _alloc_stats(self, c):
        n = self.N * c * 2
        assert self._stats_used + n <= self._stats_cap, 'stats arena too small'
        v = self._stats_arena[self._stats_used:self._stats_used + n]
        self._stats_used += n
        return v

    def _step(self, fn, launches=1, name=None):
        import sys
        if name is None:
            f = sys._getframe(1)
            name = f'{f.f_code.co_name}:{f.f_lineno}'
        self.steps.append(fn)
        self.step_names.append(name)
        self.n_launches += launches

    def _conv(self, cout, srcs, hout, wout, out, binder, **kw):
        """srcs: ops.conv_src dicts; binder(cw) packs the B operand from self.P.  Appends one launch."""
        n = kw.pop('n', self.N)
        window = any(len(s_['taps']) > 1 for s_ in srcs)
        m_rows = n * kw.get('hp', hout + 2) * kw.get('wp', wout + 2) if window else n * hout * wout
        cw = ops.ConvWeights(cout, [(s['C'], len(s['taps'])) for s in srcs], self.dev, precision=self.prec, m_rows=m_rows)
        kw.setdefault('splitk_ws', self._splitk_ws)
        desc = ops.build_conv_desc(cw, srcs, n, hout, wout, out, **kw)
        self._last_desc = desc
        self._keep.append((cw, desc, srcs, kw, out))
        if binder is not None:
            self.binders.append(lambda cw=cw: binder(cw))
        import sys
        f = sys._getframe(1)
        self._step(lambda d=desc: ops.conv_launch(d), name=f'conv {f.f_code.co_name}:{f.f_lineno} cout={cout} {hout}x{wout} srcs={[(s["C"], len(s["taps"])) for s in srcs]}')
        self.conv_flops += 2 * n * hout * wout * cout * sum(s['C'] * len(s['taps']) for s in srcs)
        self.conv_bytes += 4 * n * hout * wout * (sum(s['C'] for s in srcs) + cout * (2 if kw.get('res') is not None else 1)) + cw.bytes_per_batch
        return cw

    def _planes_of(self, act):
        """Pre-split bf16 planes of an activation for consumers that read it raw (1x1 skip convs): written by the producing conv's
        epilogue when there is one (no extra pass), else by a split kernel placed here in the plan.  None when the geometry is
        not the 2-D tiling the TMA path covers (H % 16, W % 8)."""
        # Measured on B200 (CIFAR NCSN++, batch 64, whole sampling step) with the 4-D tensor map (160-byte box rows): BF16 mode
        # 4529 -> 4658 images/s, BF16x3 mode 3802 -> 3850.  (The first tensor map had a separate 8-channel inner dimension -- 16-byte
        # rows -- and lost to the fp32 producer path in BF16x3 mode.)  On by default; DDG_ENGINE_PLANES=0 turns it off.
        on = os.environ.get('DDG_ENGINE_PLANES', '1') == '1'
        if not on or act.H % 16 != 0 or act.W % 8 != 0:
            return None
        if act.planes is None:
            act.planes = ops.alloc_planes(self.N, act.H, act.W, act.C, self.prec, self.dev)
            if act.producer is not None:
                act.producer.out_planes = act.planes.data_ptr()
            else:
                self._step(lambda a=act: ops.split_planes(a.buf, a.planes, self.prec), name='split_planes')
        return act.planes

    def _linear_rows(self, x, k, w, b, out, act_in=ops.ACT_NONE):
        """out[N][J] = act_in(x[N][k]) @ w[J][k]^T + b as one tensor-core GEMM (rows = batch samples); w is re-packed by a binder
        that must run after whatever fills w.  Falls back to the SIMT linear kernel when k is not a multiple of the K block."""
        j = w.shape[0]
        if k % ops.KB != 0:
            self._step(lambda: ops.linear(x, w, b, act_in=act_in, out=out))
            return
        self._conv(j, [ops.conv_src(x, k, ops.TAPS_1X1, act=act_in, padded=False)], 1, 1, out,
                   lambda cw: cw.pack_segment(0, w, k, k, 1, 0), out_mode=ops.OUT_NHWC, bias=b)

    def _gn_coeffs(self, acts, groups, gamma, beta, gb_stride, per_sample, eps=1e-6):
        """scale/shift [N][C] of the GroupNorm over the channel concatenation of `acts` (ncsnpp...:367 torch.cat)."""
        c = sum(a.C for a in acts)
        scale = torch.empty(self.N, c, device=self.dev)
        shift = torch.empty(self.N, c, device=self.dev)
        a = acts[0]
        b = acts[1] if len(acts) > 1 else None
        hw = a.H * a.W
        sb, cb = (b.stats, b.C) if b is not None else (None, 0)
        self._step(lambda sa=a.stats, ca=a.C, sb=sb, cb=cb, sc=scale, sh=shift: ops.gn_prepare(
            sa, ca, sb, cb, gamma, beta, gb_stride, per_sample, self.N, hw, groups, sc, sh, eps))
        self._keep.append((scale, shift))
        return scale, shift

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.P if k not in sd]
        extra = [k for k in sd if k not in self.P]
        if strict and (missing or extra):
            raise KeyError(f'state_dict mismatch: missing {missing[:4]}..., unexpected {extra[:4]}...')
        for k, buf in self.P.items():
            if k in sd and sd[k].data_ptr() != buf.data_ptr():
                buf.copy_(sd[k].detach().reshape(buf.shape))
        self.refresh()

    def refresh(self):
        """Re-derive packed operands after self.P changed (weight update): the binders' small copies, then every B-operand pack
        of the plan as one launch (recorded on the first refresh)."""
        first = self._pack_plan is None
        if first:
            self._pack_plan = ops.PackPlan()
        with ops.pack_context(self._pack_plan, 'record' if first else 'skip'):
            for b in self.binders:
                b()
        self._pack_plan.run()

    def run_steps(self):
        for s in self.steps:
            s()

    def capture(self):
        self.run_steps()  # warm-up: function attributes, lazy module loading
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run_steps()
        self.graph = g
        return g

    def replay(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run_steps()


class GeneratorEngine(_EngineBase):
    def __init__(self, cfg, batch, device='cuda', precision=3, params=None):
        cfg = arch.normalize_config(cfg)
        if not (cfg.resblock_type == 'biggan' and cfg.embedding_type == 'positional' and cfg.progressive == 'none'
                and cfg.progressive_input in ('residual', 'none') and cfg.fir and cfg.conditional
                and tuple(cfg.fir_kernel) == (1, 3, 3, 1) and cfg.num_channels_dae % 32 == 0):
            raise NotImplementedError('GeneratorEngine covers the biggan / positional / fir configuration family (README configs)')
        super().__init__(batch, device, precision, arch.ncsnpp_param_shapes(cfg), params=params)
        self.cfg = cfg
        self._build()

    def _build(self):
        cfg, N, dev, P = self.cfg, self.N, self.dev, self.P
        nf, zd = cfg.num_channels_dae, cfg.z_emb_dim
        S = cfg.image_size
        mods = arch.ncsnpp_modules(cfg)
        self.x_in = torch.zeros(N, cfg.num_channels, S, S, device=dev)
        self.t_in = torch.zeros(N, dtype=torch.int64, device=dev)
        self.z_in = torch.zeros(N, cfg.nz, device=dev)
        self.out = torch.zeros(N, cfg.num_channels, S, S, device=dev)

        # batched projection tables: AdaGN styles (layerspp.py:57) and Dense_0 (layerspp.py:298-299)
        self._style_names = [k[:-7] for k in self.shapes if k.endswith('.style.weight')]
        self._dense_names = [k[:-7] for k in self.shapes if k.endswith('Dense_0.weight')]
        self._style_off, self._dense_off = {}, {}
        jt = 0
        for nme in self._style_names:
            self._style_off[nme] = jt
            jt += self.shapes[nme + '.weight'][0]
        jd = 0
        for nme in self._dense_names:
            self._dense_off[nme] = jd
            jd += self.shapes[nme + '.weight'][0]
        self.jt, self.jd = jt, jd
        self.style_all = torch.empty(N, jt, device=dev)
        self.dense_all = torch.empty(N, jd, device=dev)
        self._w_style = torch.empty(jt, zd, device=dev); self._b_style = torch.empty(jt, device=dev)
        self._w_dense = torch.empty(jd, 4 * nf, device=dev); self._b_dense = torch.empty(jd, device=dev)

        dsts, srcs = [], []
        for nme in self._style_names:
            o = self._style_off[nme]; j = self.shapes[nme + '.weight'][0]
            dsts += [self._w_style[o:o + j], self._b_style[o:o + j]]; srcs += [P[nme + '.weight'], P[nme + '.bias']]
        for nme in self._dense_names:
            o = self._dense_off[nme]; j = self.shapes[nme + '.weight'][0]
            dsts += [self._w_dense[o:o + j], self._b_dense[o:o + j]]; srcs += [P[nme + '.weight'], P[nme + '.bias']]
        self.binders.append(lambda: torch._foreach_copy_(dsts, srcs))     # a few multi-tensor launches for the 168 slices

        used_holder = [0]
        self._step(lambda: self._stats_arena[:used_holder[0]].zero_())
        # z mapping (ncsnpp_generator_adagn.py:51-56, 271-277) and time embedding (:295-303)
        # z-mapping network and time-embedding MLP: one kernel each (ddg_mlp_rows); their last activation is applied as the
        # prologue of the consuming GEMM (style / Dense_0 projections)
        znames = ['z_transform.1'] + [f'z_transform.{3 + 2 * i}' for i in range(cfg.n_mlp)]
        zdesc = ops.make_mlp_desc([P[k + '.weight'] for k in znames], [P[k + '.bias'] for k in znames], act=ops.ACT_SILU, pixel_norm=True)
        self.zemb_pre = torch.empty(N, zd, device=dev)
        self._step(lambda: ops.mlp_rows(self.z_in, zdesc, self.zemb_pre))
        temb0 = torch.empty(N, nf, device=dev)
        self.temb = torch.empty(N, 4 * nf, device=dev)
        tdesc = ops.make_mlp_desc([P['all_modules.0.weight'], P['all_modules.1.weight']], [P['all_modules.0.bias'], P['all_modules.1.bias']],
                                  act=ops.ACT_SILU)
        self._step(lambda: ops.timestep_embedding(self.t_in, nf, out=temb0))
        self._step(lambda: ops.mlp_rows(temb0, tdesc, self.temb))
        self._linear_rows(self.zemb_pre, zd, self._w_style, self._b_style, self.style_all, act_in=ops.ACT_SILU)
        self._linear_rows(self.temb, 4 * nf, self._w_dense, self._b_dense, self.dense_all, act_in=ops.ACT_SILU)
        self._keep.append((zdesc, tdesc, temb0))

        # input image -> PNHWC (channels padded to 32); "2x-1" when data is not centered (:308-310)
        cp_in = ops.pad_c(cfg.num_channels)
        xin = Act(self, cp_in, S, S, stats=False)
        sc_in, sh_in = (1.0, 0.0) if cfg.centered else (2.0, -1.0)
        self._step(lambda: ops.to_pnhwc(self.x_in, cpad=cp_in, out=xin.buf, scale=sc_in, shift=sh_in))

        it = iter(mods[2:])
        m = next(it)
        assert m['kind'] == 'conv3'
        nme = f"all_modules.{m['idx']}"
        h0 = Act(self, nf, S, S)
        self._conv(nf, [ops.conv_src(xin.buf, cp_in, ops.TAPS_3X3)], S, S, h0.buf,
                   lambda cw, nme=nme: cw.pack_segment(0, P[nme + '.weight'], cfg.num_channels, cfg.num_channels * 9, 9, 1),
                   bias=P[nme + '.bias'], stats=h0.stats)
        h0.producer = self._last_desc
        hs = [h0]
        pyramid = xin
        h = h0
        n_down = (len(cfg.ch_mult) - 1)
        # walk the module list; the skip stack discipline follows ncsnpp_generator_adagn.py:317-413
        mods_rest = list(it)
        i = 0
        nres = len(cfg.ch_mult)
        # ---- down path ----
        for lvl in range(nres):
            for _ in range(cfg.num_res_blocks):
                m = mods_rest[i]; i += 1
                h = self._resblock(m, [hs[-1]])
                if m['res'] in cfg.attn_resolutions:
                    h = self._attn(mods_rest[i], h); i += 1
                hs.append(h)
            if lvl != nres - 1:
                m = mods_rest[i]; i += 1
                assert m['down']
                h = self._resblock(m, [hs[-1]])
                if cfg.progressive_input == 'residual':
                    m = mods_rest[i]; i += 1
                    assert m['kind'] == 'pyrdown'
                    h = self._pyramid_down(m, pyramid, h)
                    pyramid = h
                hs.append(h)
        # ---- middle ----
        h = hs[-1]
        h = self._resblock(mods_rest[i], [h]); i += 1
        h = self._attn(mods_rest[i], h); i += 1
        h = self._resblock(mods_rest[i], [h]); i += 1
        # ---- up path ----
        for lvl in reversed(range(nres)):
            for _ in range(cfg.num_res_blocks + 1):
                h = self._resblock(mods_rest[i], [h, hs.pop()]); i += 1
            if (S // 2 ** lvl) in cfg.attn_resolutions:
                h = self._attn(mods_rest[i], h); i += 1
            if lvl != 0:
                assert mods_rest[i]['up']
                h = self._resblock(mods_rest[i], [h]); i += 1
        assert not hs and n_down >= 0
        # ---- head: GroupNorm(affine) -> SiLU -> conv3x3 -> tanh (:420-431) ----
        g, c = mods_rest[i], mods_rest[i + 1]
        assert g['kind'] == 'gn' and c['kind'] == 'conv3' and i + 2 == len(mods_rest)
        gname, cname = f"all_modules.{g['idx']}", f"all_modules.{c['idx']}"
        sc, sh = self._gn_coeffs([h], _groups(h.C), P[gname + '.weight'], P[gname + '.bias'], 0, 0)
        self._conv(cfg.num_channels, [ops.conv_src(h.buf, h.C, ops.TAPS_3X3, scale=sc, shift=sh, act=ops.ACT_SILU)], S, S,
                   self.out, lambda cw: cw.pack_conv_weight(0, P[cname + '.weight']), out_mode=ops.OUT_NCHW,
                   bias=P[cname + '.bias'], out_act=ops.ACT_NONE if cfg.not_use_tanh else ops.ACT_TANH)
        used_holder[0] = self._stats_used

    def _adagn(self, acts, prefix):
        c = sum(a.C for a in acts)
        off = self._style_off[prefix + '.style']
        base = self.style_all.data_ptr()
        return self._gn_coeffs(acts, _groups(c), base + 4 * off, base + 4 * (off + c), self.jt, 1)

    def _resblock(self, m, X):
        """ResnetBlockBigGANpp_Adagn.forward (layerspp.py:278-310): [gn_prepare] [fir, fir] conv0 [gn_prepare] conv1."""
        cfg, Pm = self.cfg, self.P
        assert m['kind'] == 'res'
        Pn = f"all_modules.{m['idx']}."
        cin, out_ch, up, down = m['i'], m['o'], m['up'], m['down']
        assert cin == sum(a.C for a in X)
        H, W = X[0].H, X[0].W
        sc0, sh0 = self._adagn(X, Pn + 'GroupNorm_0')
        has_skip_conv = (cin != out_ch) or up or down
        if up or down:
            xsrc = X[0]
            H2, W2 = (2 * H, 2 * W) if up else (H // 2, W // 2)
            hf = Act(self, cin, H2, W2, stats=False)
            xf = Act(self, cin, H2, W2, stats=False)
            mode = 1 if up else 2
            # NB: bind through default arguments -- closures capture variables, not values
            self._step(lambda i=xsrc.buf, o=hf.buf, m_=mode, a_=sc0, b_=sh0: ops.fir_pnhwc(i, m_, o, a_, b_, ops.ACT_SILU))
            self._step(lambda i=xsrc.buf, o=xf.buf, m_=mode: ops.fir_pnhwc(i, m_, o))
            src0 = [ops.conv_src(hf.buf, cin, ops.TAPS_3X3)]
            skip = [xf]
            H, W = H2, W2
        else:
            src0, off = [], 0
            for xa in X:
                src0.append(ops.conv_src(xa.buf, xa.C, ops.TAPS_3X3, scale=sc0.data_ptr() + 4 * off,
                                         shift=sh0.data_ptr() + 4 * off, act=ops.ACT_SILU, ss_stride=cin))
                off += xa.C
            skip = X
        h1 = Act(self, out_ch, H, W)

        def bind0(cw, segs=[s['C'] for s in src0]):
            c0 = 0
            for i, c in enumerate(segs):
                cw.pack_segment(i, Pm[Pn + 'Conv_0.weight'], c, cin * 9, 9, 1, elem_offset=c0 * 9)
                c0 += c
        doff = self._dense_off[Pn + 'Dense_0']
        self._conv(out_ch, src0, H, W, h1.buf, bind0, bias=Pm[Pn + 'Conv_0.bias'], stats=h1.stats,
                   addvec=self.dense_all.data_ptr() + 4 * doff, addvec_stride=self.jd)
        sc1, sh1 = self._adagn([h1], Pn + 'GroupNorm_1')
        out = Act(self, out_ch, H, W)
        src1 = [ops.conv_src(h1.buf, out_ch, ops.TAPS_3X3, scale=sc1, shift=sh1, act=ops.ACT_SILU)]
        scale = RSQRT2 if cfg.skip_rescale else 1.0
        if has_skip_conv:
            # K segments: the 3x3 over the normalised h1, then the 1x1 skip conv over the raw block input(s).  (Measured: putting the
            # 1-tap segments first, so that they would be produced under the previous tile's 3x3 MMAs, is 2 % slower end to end.)
            skip_first = os.environ.get('DDG_SKIP_SEG_FIRST') is not None
            skip_srcs = [ops.conv_src(xs.buf, xs.C, ops.TAPS_1X1, planes=self._planes_of(xs)) for xs in skip]
            src1 = skip_srcs + src1 if skip_first else src1 + skip_srcs
            i3 = len(skip_srcs) if skip_first else 0
            i1 = 0 if skip_first else 1
            bias_buf = torch.empty(out_ch, device=self.dev)

            def bind1(cw, segs=[xs.C for xs in skip]):
                cw.pack_conv_weight(i3, Pm[Pn + 'Conv_1.weight'])
                c0 = 0
                for i, c in enumerate(segs):
                    cw.pack_segment(i1 + i, Pm[Pn + 'Conv_2.weight'], c, cin, 1, 0, elem_offset=c0)
                    c0 += c
                torch.add(Pm[Pn + 'Conv_1.bias'], Pm[Pn + 'Conv_2.bias'], out=bias_buf)
            self._conv(out_ch, src1, H, W, out.buf, bind1, bias=bias_buf, out_scale=scale, stats=out.stats)
        else:
            assert len(X) == 1
            self._conv(out_ch, src1, H, W, out.buf, lambda cw: cw.pack_conv_weight(0, Pm[Pn + 'Conv_1.weight']),
                       bias=Pm[Pn + 'Conv_1.bias'], res=X[0].buf, out_scale=scale, stats=out.stats)
        out.producer = self._last_desc
        return out

    def _attn(self, m, X):
        """AttnBlockpp.forward (layerspp.py:108-124): GN -> fused QKV 1x1 -> batched QK^T -> softmax -> batched PV ->
        NIN_3 + residual / sqrt(2)."""
        assert m['kind'] == 'attn' and m['c'] == X.C
        Pm = self.P
        Pn = f"all_modules.{m['idx']}."
        N, C, H, W = self.N, X.C, X.H, X.W
        T = H * W
        Tp = ops.pad_c(T)
        sc, sh = self._gn_coeffs([X], _groups(C), Pm[Pn + 'GroupNorm_0.weight'], Pm[Pn + 'GroupNorm_0.bias'], 0, 0)
        qkv = torch.zeros(N, T, 3 * C, device=self.dev)
        wqkv = torch.empty(C, 3 * C, device=self.dev)
        bqkv = torch.empty(3 * C, device=self.dev)

        def bind_qkv(cw):
            torch.cat([Pm[Pn + f'NIN_{j}.W'] for j in range(3)], 1, out=wqkv)
            torch.cat([Pm[Pn + f'NIN_{j}.b'] for j in range(3)], 0, out=bqkv)
            cw.pack_nin_weight(0, wqkv)
        self._conv(3 * C, [ops.conv_src(X.buf, C, ops.TAPS_1X1, scale=sc, shift=sh)], H, W, qkv, bind_qkv,
                   out_mode=ops.OUT_NHWC, bias=bqkv)
        if T == 256 and C == 256 and not getattr(self, 'unfused_attention', False):
            # fused core (ddg_attention_fwd): QK^T -> softmax -> PV -> NIN_3 + residual in one kernel per (sample, 128 queries);
            # logits / weights / attention output never leave the SM and there are no per-image operand packs
            w3 = ops.ConvWeights(C, [(C, 1)], self.dev, precision=self.prec, nt=256)
            self.binders.append(lambda: w3.pack_nin_weight(0, Pm[Pn + 'NIN_3.W']))
            out = Act(self, C, H, W)
            d = ops.attention_desc(qkv, w3, Pm[Pn + 'NIN_3.b'], X.buf, out.buf, out.stats, N, H, W, C,
                                   RSQRT2 if self.cfg.skip_rescale else 1.0, precision=self.prec)
            self._step(lambda d=d: ops.attention_launch(d), name=f'attn_fused {H}x{W} C={C}')
            self.conv_flops += 2 * N * T * T * C * 2 + 2 * N * T * C * C
            self._keep.append((w3, d, qkv, wqkv, bqkv))
            return out
        # per-image K operand: B[co = key t][ci = channel]
        wk = ops.ConvWeights(T, [(C, 1)], self.dev, precision=self.prec, batch=N, m_rows=N * T)
        self._step(lambda: wk.pack_segment(0, qkv, C, 3 * C, 1, 0, w_batch_stride=T * 3 * C, elem_offset=C))
        s = torch.zeros(N, T, Tp, device=self.dev)
        d = ops.build_conv_desc(wk, [ops.conv_src(qkv, C, ops.TAPS_1X1, padded=False, pitch=3 * C)], N, 1, T, s,
                                out_mode=ops.OUT_NHWC, out_c=Tp, out_scale=float(C) ** -0.5, batch_rows=T)
        self._step(lambda d=d: ops.conv_launch(d))
        p = torch.zeros(N, T, Tp, device=self.dev)
        self._step(lambda: ops.softmax_rows(s, p, N * T, T, Tp, Tp))
        # per-image V operand: B[co = channel][ci = key t] = v[t][c]
        wv = ops.ConvWeights(C, [(Tp, 1)], self.dev, precision=self.prec, batch=N, m_rows=N * T)
        self._step(lambda: wv.pack_segment(0, qkv, T, 1, 3 * C, 0, w_batch_stride=T * 3 * C, elem_offset=2 * C))
        o = torch.zeros(N, T, C, device=self.dev)
        d2 = ops.build_conv_desc(wv, [ops.conv_src(p, Tp, ops.TAPS_1X1, padded=False)], N, 1, T, o, out_mode=ops.OUT_NHWC,
                                 batch_rows=T)
        self._step(lambda d=d2: ops.conv_launch(d))
        self.conv_flops += 2 * N * T * T * C * 2
        self._keep.append((wk, wv, d, d2, s, p, o, qkv, wqkv, bqkv))
        out = Act(self, C, H, W)
        self._conv(C, [ops.conv_src(o, C, ops.TAPS_1X1, padded=False)], H, W, out.buf,
                   lambda cw: cw.pack_nin_weight(0, Pm[Pn + 'NIN_3.W']), bias=Pm[Pn + 'NIN_3.b'], res=X.buf,
                   out_scale=RSQRT2 if self.cfg.skip_rescale else 1.0, stats=out.stats)
        out.producer = self._last_desc
        return out

    def _pyramid_down(self, m, pyr, h):
        """layerspp.Downsample(fir, with_conv) = conv_downsample_2d + bias (up_or_down_sampling.py:149-183, :52-59) then the
        progressive-input residual (ncsnpp_generator_adagn.py:343-350): FIR pad(2,2) -> space-to-depth -> 2x2-tap stride-1
        conv over 4*C channels, with the residual and 1/sqrt(2) in the epilogue."""
        Pm = self.P
        Pn = f"all_modules.{m['idx']}.Conv2d_0"
        N = self.N
        cp = pyr.C
        cin_real = m['i']
        Ho, Wo = pyr.H // 2, pyr.W // 2
        cout = h.C
        assert cout == m['o'] and h.H == Ho
        s2d = torch.zeros(N, Ho + 3, Wo + 3, 4 * cp, device=self.dev)
        self._step(lambda i=pyr.buf, o=s2d: ops.fir_pnhwc(i, 3, o))
        out = Act(self, cout, Ho, Wo)
        w2 = torch.zeros(cout, 2, 2, cp, 2, 2, device=self.dev)

        def bind(cw):
            ops.s2d_weights(Pm[Pn + '.weight'], cout, cin_real, cp, out=w2)       # [Cout, Cin, 3, 3] -> [Cout, 2, 2, cp, 2, 2]
            cw.pack_segment(0, w2, 4 * cp, 4 * cp * 4, 4, 1)
        scale = RSQRT2 if self.cfg.skip_rescale else 1.0
        self._conv(cout, [ops.conv_src(s2d, 4 * cp, ops.TAPS_2X2)], Ho, Wo, out.buf, bind, hp=Ho + 3, wp=Wo + 3,
                   bias=Pm[Pn + '.bias'], res=h.buf, out_scale=scale, stats=out.stats)
        out.producer = self._last_desc
        self._keep.append((s2d, w2))
        return out

    def forward(self, x, t, z):
        """x [N,C,H,W] fp32, t [N] int64, z [N,nz] -> static output buffer [N,C,H,W] (overwritten by the next call)."""
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        self.z_in.copy_(z)
        self.replay()
        return self.out


class DiscriminatorEngine(_EngineBase):
    """Discriminator_small / Discriminator_large forward (discriminator.py:134-167 / :205-238), act = LeakyReLU(0.2)."""

    def __init__(self, nc, ngf, t_emb_dim, image_size, batch, large=False, device='cuda', precision=3, params=None):
        super().__init__(batch, device, precision, arch.discriminator_param_shapes(nc, ngf, t_emb_dim, large), params=params)
        self.nc, self.ngf, self.t_emb_dim, self.S, self.large = nc, ngf, t_emb_dim, image_size, large
        # narrow maps (2*ngf, 4*ngf not multiples of 32) live in buffers padded to 32 channels whose extra channels stay zero;
        # the minibatch-stddev / final_conv stage needs its 8*ngf channels unpadded
        if nc % 2 != 0 or (8 * ngf) % 32 != 0:
            raise NotImplementedError('DiscriminatorEngine needs an even nc and ngf % 4 == 0')
        self._build()

    def _build(self):
        N, dev, P, S = self.N, self.dev, self.P, self.S
        te_dim = self.t_emb_dim
        blocks = arch.discriminator_blocks(self.ngf, self.large)
        self.x_in = torch.zeros(N, self.nc // 2, S, S, device=dev)
        self.xt_in = torch.zeros(N, self.nc // 2, S, S, device=dev)
        self.t_in = torch.zeros(N, dtype=torch.int64, device=dev)
        self.out = torch.zeros(N, 1, device=dev)
        # time embedding (discriminator.py:19-36, :135); consumers apply the outer LeakyReLU on load
        te0 = torch.empty(N, te_dim, device=dev); te1 = torch.empty(N, te_dim, device=dev); te = torch.empty(N, te_dim, device=dev)
        self._step(lambda: ops.timestep_embedding(self.t_in, te_dim, out=te0))
        self._step(lambda: ops.linear(te0, P['t_embed.main.0.weight'], P['t_embed.main.0.bias'], out=te1))
        self._step(lambda: ops.linear(te1, P['t_embed.main.2.weight'], P['t_embed.main.2.bias'], act_in=ops.ACT_LEAKY, out=te))
        jd = sum(b for _, b, _ in blocks)
        self.jd = jd
        self.dense_all = torch.empty(N, jd, device=dev)
        w_dense = torch.empty(jd, te_dim, device=dev); b_dense = torch.empty(jd, device=dev)
        offs = []
        o = 0
        for _, b, _ in blocks:
            offs.append(o); o += b

        def bind_dense():
            for i, (_, b, _) in enumerate(blocks):
                w_dense[offs[i]:offs[i] + b].copy_(P[f'conv{i + 1}.dense_t1.weight'])
                b_dense[offs[i]:offs[i] + b].copy_(P[f'conv{i + 1}.dense_t1.bias'])
        self.binders.append(bind_dense)
        self._linear_rows(te, te_dim, w_dense, b_dense, self.dense_all, act_in=ops.ACT_LEAKY)
        self._keep.append((te0, te1, te, w_dense, b_dense))
        # input: cat(x, x_t) (:138) -> PNHWC
        cp_in = ops.pad_c(self.nc)
        xin = Act(self, cp_in, S, S, stats=False)
        self._step(lambda: ops.to_pnhwc(self.x_in, self.xt_in, cpad=cp_in, out=xin.buf))
        c0 = 2 * self.ngf
        h = Act(self, ops.pad_c(c0), S, S, stats=False)
        self._conv(c0, [ops.conv_src(xin.buf, cp_in, ops.TAPS_1X1)], S, S, h.buf,
                   lambda cw: cw.pack_segment(0, P['start_conv.weight'], self.nc, self.nc, 1, 0), bias=P['start_conv.bias'],
                   out_c=h.C)
        self.stage_acts = [h]                     # start_conv and block outputs (PNHWC), for per-stage parity checks
        for i, (a, b, ds) in enumerate(blocks):
            h = self._down_block(i, h, a, b, ds, offs[i])
            self.stage_acts.append(h)
        # minibatch stddev (:150-158) as a 1-real-channel PNHWC tensor, then final conv over 512 + 1 channels
        C, H, W = h.C, h.H, h.W
        group = min(N, 4)
        assert N % group == 0, 'batch must be divisible by the stddev group (as in the reference view())'
        sdv = Act(self, 32, H, W, stats=False)
        self._step(lambda i=h.buf, o=sdv.buf: ops.minibatch_stddev(i, o, group))
        f = Act(self, C, H, W, stats=False)

        def bind_final(cw):
            w = P['final_conv.weight']  # [C, C+1, 3, 3]
            cw.pack_segment(0, w, C, (C + 1) * 9, 9, 1)
            cw.pack_segment(1, w, 1, (C + 1) * 9, 9, 1, elem_offset=C * 9)
        self._conv(C, [ops.conv_src(h.buf, C, ops.TAPS_3X3), ops.conv_src(sdv.buf, 32, ops.TAPS_3X3)], H, W, f.buf, bind_final,
                   bias=P['final_conv.bias'])
        pooled = torch.empty(N, C, device=dev)
        self.final_feat, self.pooled = f, pooled   # final_conv output before the activation; pooled = sum(leaky(f))
        self._step(lambda i=f.buf: ops.spatial_sum(i, ops.ACT_LEAKY, out=pooled))
        self._step(lambda: ops.linear(pooled, P['end_linear.weight'], P['end_linear.bias'], out=self.out))
        self._keep.append((pooled,))

    def _down_block(self, i, X, cin, cout, ds, doff):
        """DownConvBlock.forward (discriminator.py:76-94).  cin / cout are the real channel counts; buffers carry pad_c() channels
        (zero beyond the real ones, zero weight columns for them)."""
        P = self.P
        Pn = f'conv{i + 1}.'
        H, W = X.H, X.W
        cip, cop = ops.pad_c(cin), ops.pad_c(cout)
        assert X.C == cip
        h1 = Act(self, cop, H, W, stats=False)
        self._conv(cout, [ops.conv_src(X.buf, cip, ops.TAPS_3X3, act=ops.ACT_LEAKY)], H, W, h1.buf,
                   lambda cw: cw.pack_conv_weight(0, P[Pn + 'conv1.0.weight']), bias=P[Pn + 'conv1.0.bias'],
                   addvec=self.dense_all.data_ptr() + 4 * doff, addvec_stride=self.jd, out_c=cop)
        if ds:
            H2, W2 = H // 2, W // 2
            hf = Act(self, cop, H2, W2, stats=False)
            xf = Act(self, cip, H2, W2, stats=False)
            self._step(lambda i=h1.buf, o=hf.buf: ops.fir_pnhwc(i, 2, o, None, None, ops.ACT_LEAKY))
            self._step(lambda i=X.buf, o=xf.buf: ops.fir_pnhwc(i, 2, o))
            srcs = [ops.conv_src(hf.buf, cop, ops.TAPS_3X3), ops.conv_src(xf.buf, cip, ops.TAPS_1X1)]
            H, W = H2, W2
        else:
            srcs = [ops.conv_src(h1.buf, cop, ops.TAPS_3X3, act=ops.ACT_LEAKY), ops.conv_src(X.buf, cip, ops.TAPS_1X1)]
        out = Act(self, cop, H, W, stats=False)

        def bind(cw):
            cw.pack_conv_weight(0, P[Pn + 'conv2.0.weight'])
            cw.pack_segment(1, P[Pn + 'skip.0.weight'], cin, cin, 1, 0)
        self._conv(cout, srcs, H, W, out.buf, bind, bias=P[Pn + 'conv2.0.bias'], out_scale=RSQRT2, out_c=cop)
        return out

    def forward(self, x, t, x_t):
        self.x_in.copy_(x)
        self.xt_in.copy_(x_t)
        self.t_in.copy_(t)
        self.replay()
        return self.out
