"""GPU parity of every C-ABI kernel against the CPU oracle (bit-for-tolerance: fp32 path <= 1e-4 relative L2,
north_star).  All calls go through libddgan_b200.so via ddgan_b200.ops."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ddgan_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4  # north_star: FP32 mode within 1e-4 relative L2 per op
DEV = 'cuda'


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.fixture(scope='module')
def ops():
    from ddgan_b200 import ops as _ops
    return _ops


def test_upfirdn2d_golden_cases(ops, golden):
    for c in golden['upfirdn2d_cases']:
        x = seeded(c['shape'], c['seed'])
        n, ch, h, w = x.shape
        y = ops.upfirdn2d_raw(x.view(n * ch, h, w).to(DEV), c['kernel'].to(DEV), c['up'], c['up'], c['down'], c['down'],
                              c['pad'][0], c['pad'][1], c['pad'][0], c['pad'][1])
        ref = c['out']
        assert tuple(y.shape[1:]) == tuple(ref.shape[2:]), c['name']
        assert O.rel_l2(y.view(ref.shape).cpu(), ref) < 1e-6, c['name']


@pytest.mark.parametrize('shape,up,down,pad', [((64, 128, 32, 32), 1, 2, (1, 1)), ((64, 256, 16, 16), 2, 1, (2, 1)),
                                               ((8, 3, 32, 32), 1, 1, (2, 2)), ((2, 4, 33, 31), 2, 1, (2, 1)),
                                               ((1, 1, 1, 1), 2, 1, (2, 1)), ((2, 3, 5, 7), 1, 2, (1, 1))])
def test_upfirdn2d_cifar_shapes(ops, shape, up, down, pad):
    x = seeded(shape, 1)
    k = torch.from_numpy(O.setup_fir_kernel([1, 3, 3, 1])) * (up ** 2)
    n, c, h, w = shape
    y = ops.upfirdn2d_raw(x.view(n * c, h, w).to(DEV), k.to(DEV), up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    ref = O.upfirdn2d(x, k, up, down, pad)
    assert O.rel_l2(y.view(ref.shape).cpu(), ref) < 1e-6


def test_fused_bias_act(ops):
    x = seeded((4, 16, 8, 8), 2); b = seeded((16,), 3)
    y = ops.fused_bias_act(x.to(DEV), b.to(DEV), None, 3, 0, 0.2, 2 ** 0.5)
    assert O.rel_l2(y.cpu(), O.fused_leaky_relu(x, b)) < 1e-6
    # other slope, odd spatial size (scalar path)
    x = seeded((3, 5, 3, 3), 4); b = seeded((5,), 5)
    y = ops.fused_bias_act(x.to(DEV), b.to(DEV), None, 3, 0, 0.1, 1.5)
    assert O.rel_l2(y.cpu(), O.fused_leaky_relu(x, b, 0.1, 1.5)) < 1e-6
    # gradient mode (fused_act.py:28-50)
    out = O.fused_leaky_relu(x, b, 0.1, 1.5); g = seeded(tuple(x.shape), 6)
    gi = ops.fused_bias_act(g.to(DEV), None, out.to(DEV), 3, 1, 0.1, 1.5)
    gref, gb = O.fused_leaky_relu_grad(g, out, 0.1, 1.5)
    assert O.rel_l2(gi.cpu(), gref) < 1e-6
    assert O.rel_l2(ops.channel_sum(gi).cpu(), gb) < 1e-5
    # empty input
    assert ops.fused_bias_act(torch.empty(0, 4, device=DEV), None, None, 3, 0, 0.2, 1.0).numel() == 0


@pytest.mark.parametrize('n,c,h,act,ada', [(4, 128, 32, 1, True), (2, 256, 16, 0, False), (3, 384, 8, 1, True),
                                           (2, 12, 5, 0, True), (2, 64, 64, 1, True),
                                           # slabs split over thread-block clusters: 256 KB -> 8 CTAs, 1 MB -> 16 CTAs
                                           (2, 64, 128, 1, True), (2, 128, 128, 0, False), (1, 64, 256, 1, True)])
def test_groupnorm_fwd_bwd(ops, n, c, h, act, ada):
    x = seeded((n, c, h, h), 7) * 2 + 0.5
    G = O.num_groups(c)
    if ada:
        gamma = 1 + seeded((n, c), 8) * 0.2; beta = seeded((n, c), 9) * 0.2
    else:
        gamma = 1 + seeded((c,), 8) * 0.2; beta = seeded((c,), 9) * 0.2
    xr = x.clone().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    y0 = O.group_norm(xr, G)
    if ada:
        ref = gr[:, :, None, None] * y0 + br[:, :, None, None]
    else:
        ref = gr.view(1, -1, 1, 1) * y0 + br.view(1, -1, 1, 1)
    if act == 1:
        ref = F.silu(ref)
    y, mean, rstd = ops.groupnorm_fwd(x.to(DEV), G, gamma.to(DEV), beta.to(DEV), per_sample=ada, act=act)
    assert O.rel_l2(y.cpu(), ref) < 1e-5
    dy = seeded(tuple(x.shape), 10)
    ref.backward(dy)
    dx, dg, db = ops.groupnorm_bwd(x.to(DEV), dy.to(DEV), G, mean, rstd, gamma.to(DEV), beta.to(DEV), per_sample=ada, act=act)
    assert O.rel_l2(dx.cpu(), xr.grad) < TOL
    if ada:
        assert O.rel_l2(dg.cpu(), gr.grad) < TOL and O.rel_l2(db.cpu(), br.grad) < TOL
    else:
        assert O.rel_l2(dg.sum(0).cpu(), gr.grad) < TOL and O.rel_l2(db.sum(0).cpu(), br.grad) < TOL


def test_small_kernels(ops):
    t = torch.tensor([0, 1, 2, 3, 3, 0])
    for dim in (128, 256, 32):
        e = ops.timestep_embedding(t.to(DEV), dim)
        assert O.rel_l2(e.cpu(), O.timestep_embedding(t, dim)) < 1e-6
    x = seeded((70, 100), 11); W = seeded((256, 100), 12, 0.1); b = seeded((256,), 13)
    zn = x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)
    y = ops.linear(x.to(DEV), W.to(DEV), b.to(DEV), act_out=ops.ACT_SILU, pixel_norm=True)
    assert O.rel_l2(y.cpu(), F.silu(F.linear(zn, W, b))) < 1e-5
    x = seeded((64, 512), 14); W = seeded((300, 512), 15, 0.05)
    y = ops.linear(x.to(DEV), W.to(DEV), None, act_in=ops.ACT_SILU)
    assert O.rel_l2(y.cpu(), F.linear(F.silu(x), W)) < 1e-5
    cfg = O.cifar10_config()
    co, pc = O.diffusion_coefficients(cfg), O.posterior_coefficients(cfg)
    x0 = seeded((6, 3, 32, 32), 16); n0 = seeded((6, 3, 32, 32), 17); n1 = seeded((6, 3, 32, 32), 18)
    xt, xtp1 = ops.q_sample_pairs(x0.to(DEV), n0.to(DEV), n1.to(DEV), t.to(DEV), co.a_s_cum.to(DEV), co.sigmas_cum.to(DEV),
                                  co.a_s.to(DEV), co.sigmas.to(DEV))
    rt, rtp1 = O.q_sample_pairs(co, x0, t, n0, n1)
    assert O.rel_l2(xt.cpu(), rt) < 1e-6 and O.rel_l2(xtp1.cpu(), rtp1) < 1e-6
    xp = ops.sample_posterior(x0.to(DEV), n0.to(DEV), n1.to(DEV), t.to(DEV), pc.posterior_mean_coef1.to(DEV),
                              pc.posterior_mean_coef2.to(DEV), pc.posterior_log_variance_clipped.to(DEV))
    assert O.rel_l2(xp.cpu(), O.sample_posterior(pc, x0, n0, t, n1)) < 1e-6
    # ragged (non multiple of 4) size takes the scalar path
    x0 = seeded((3, 1, 5, 5), 19); n0 = seeded((3, 1, 5, 5), 20); n1 = seeded((3, 1, 5, 5), 21); t3 = torch.tensor([0, 2, 3])
    xp = ops.sample_posterior(x0.to(DEV), n0.to(DEV), n1.to(DEV), t3.to(DEV), pc.posterior_mean_coef1.to(DEV),
                              pc.posterior_mean_coef2.to(DEV), pc.posterior_log_variance_clipped.to(DEV))
    assert O.rel_l2(xp.cpu(), O.sample_posterior(pc, x0, n0, t3, n1)) < 1e-6


def test_mlp_rows_single_kernel(ops):
    """z-mapping network (PixelNorm + 5 dense + SiLU between) and a 2-layer time-embedding MLP as one launch each."""
    n, nz, zd = 13, 100, 256                                   # rows not a multiple of the 8-row CTA, K0 not a multiple of 64
    z = seeded((n, nz), 70)
    Ws = [seeded((zd, nz), 71, 0.1)] + [seeded((zd, zd), 72 + i, 0.08) for i in range(4)]
    bs = [seeded((zd,), 80 + i, 0.1) for i in range(5)]
    h = z / torch.sqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i, (w, b) in enumerate(zip(Ws, bs)):
        h = F.linear(h, w, b)
        if i < 4:
            h = F.silu(h)
    desc = ops.make_mlp_desc([w.to(DEV) for w in Ws], [b.to(DEV) for b in bs], act=ops.ACT_SILU, pixel_norm=True)
    keep = desc  # noqa: F841 (device tensors above stay alive through the local lists below)
    Wd = [w.to(DEV) for w in Ws]; bd = [b.to(DEV) for b in bs]
    desc = ops.make_mlp_desc(Wd, bd, act=ops.ACT_SILU, pixel_norm=True)
    out = ops.mlp_rows(z.to(DEV), desc, torch.empty(n, zd, device=DEV))
    assert O.rel_l2(out.cpu(), h) < 1e-5
    # 128 -> 512 -> 512 (temb MLP shape), no normalisation, 64 rows
    t0 = seeded((64, 128), 90); W0 = seeded((512, 128), 91, 0.1).to(DEV); W1 = seeded((512, 512), 92, 0.05).to(DEV)
    b0 = seeded((512,), 93, 0.1).to(DEV); b1 = seeded((512,), 94, 0.1).to(DEV)
    ref = F.linear(F.silu(F.linear(t0, W0.cpu(), b0.cpu())), W1.cpu(), b1.cpu())
    out = ops.mlp_rows(t0.to(DEV), ops.make_mlp_desc([W0, W1], [b0, b1]), torch.empty(64, 512, device=DEV))
    assert O.rel_l2(out.cpu(), ref) < 1e-5


def test_layout_roundtrip_and_fir(ops):
    a = seeded((3, 3, 8, 8), 22); b = seeded((3, 3, 8, 8), 23)
    p = ops.to_pnhwc(a.to(DEV), b.to(DEV), cpad=32)
    assert p.shape == (3, 10, 10, 32)
    back = ops.from_pnhwc(p, 6)
    assert torch.equal(back.cpu(), torch.cat([a, b], 1))
    assert float(p[:, 0].abs().max()) == 0 and float(p[..., 6:].abs().max()) == 0
    # FIR on PNHWC with fused affine + SiLU
    x = seeded((2, 32, 8, 8), 24); sc = 0.5 + torch.rand(2, 32, generator=torch.Generator().manual_seed(1)); sh = seeded((2, 32), 25, 0.3)
    xp = ops.to_pnhwc(x.to(DEV))
    tr = F.silu(x * sc[:, :, None, None] + sh[:, :, None, None])
    up = ops.fir_pnhwc(xp, 1, ops.alloc_pnhwc(2, 16, 16, 32, DEV), sc.to(DEV), sh.to(DEV), ops.ACT_SILU)
    assert O.rel_l2(ops.from_pnhwc(up).cpu(), O.upsample_2d(tr)) < 1e-5
    dn = ops.fir_pnhwc(xp, 2, ops.alloc_pnhwc(2, 4, 4, 32, DEV))
    assert O.rel_l2(ops.from_pnhwc(dn).cpu(), O.downsample_2d(x)) < 1e-5
    # several tiles per image with partial edge tiles (24 -> 48 = 3 x 16, 24 -> 12 = 8 + 4) and two 32-channel blocks
    x2 = seeded((2, 64, 24, 24), 26); sc2 = 0.5 + torch.rand(2, 64, generator=torch.Generator().manual_seed(3)); sh2 = seeded((2, 64), 27, 0.3)
    xp2 = ops.to_pnhwc(x2.to(DEV))
    tr2 = F.silu(x2 * sc2[:, :, None, None] + sh2[:, :, None, None])
    up2 = ops.fir_pnhwc(xp2, 1, ops.alloc_pnhwc(2, 48, 48, 64, DEV), sc2.to(DEV), sh2.to(DEV), ops.ACT_SILU)
    assert O.rel_l2(ops.from_pnhwc(up2).cpu(), O.upsample_2d(tr2)) < 1e-5
    assert float(up2[:, 0].abs().max()) == 0 and float(up2[:, :, -1].abs().max()) == 0
    dn2 = ops.fir_pnhwc(xp2, 2, ops.alloc_pnhwc(2, 12, 12, 64, DEV), sc2.to(DEV), sh2.to(DEV), ops.ACT_SILU)
    assert O.rel_l2(ops.from_pnhwc(dn2).cpu(), O.downsample_2d(tr2)) < 1e-5
    # mode 3: pad(2,2) FIR stored space-to-depth: cell (i,j), block (py,px) = fir[2i+py][2j+px]
    s2d = ops.fir_pnhwc(xp, 3, torch.zeros(2, 4 + 3, 4 + 3, 128, device=DEV))
    k = torch.from_numpy(O.setup_fir_kernel([1, 3, 3, 1]))
    fir = O.upfirdn2d(x, k, pad=(2, 2))  # [2, 32, 9, 9]
    firp = F.pad(fir, (0, 1, 0, 1))      # 10 x 10, zero at index 9
    cells = firp.view(2, 32, 5, 2, 5, 2).permute(0, 2, 4, 3, 5, 1).reshape(2, 5, 5, 128)
    assert O.rel_l2(s2d[:, 1:6, 1:6, :].cpu(), cells) < 1e-5


def _conv_case(ops, n, cin, cout, h, k, prec=3, affine=False, act=0, res=False, nchw=False, msub=0, temb=False, skip1x1=0, force_linear=0, splitk=None):
    x = seeded((n, cin, h, h), 30)
    w = seeded((cout, cin, k, k), 31) / math.sqrt(cin * k * k)
    b = seeded((cout,), 32, 0.1)
    xin = x
    sc = sh = None
    if affine:
        sc = torch.rand(n, cin, generator=torch.Generator().manual_seed(2)) + 0.5; sh = seeded((n, cin), 33, 0.3)
        xin = x * sc[:, :, None, None] + sh[:, :, None, None]
    if act == 1:
        xin = F.silu(xin)
    elif act == 2:
        xin = F.leaky_relu(xin, 0.2)
    ref = F.conv2d(xin, w, b, padding=k // 2)
    tv = None
    if temb:
        tv = seeded((n, cout), 34)
        ref = ref + tv[:, :, None, None]
    cp = ops.pad_c(cin)
    taps = ops.TAPS_3X3 if k == 3 else ops.TAPS_1X1
    segs = [(cp, len(taps))]
    srcs = []
    xs = w2 = None
    if skip1x1:
        xs = seeded((n, skip1x1, h, h), 35); w2 = seeded((cout, skip1x1, 1, 1), 36) / math.sqrt(skip1x1)
        ref = ref + F.conv2d(xs, w2)
        segs.append((ops.pad_c(skip1x1), 1))
    r = None
    if res:
        r = seeded((n, cout, h, h), 37)
        ref = (ref + r) / math.sqrt(2)
    cw = ops.ConvWeights(cout, segs, DEV, precision=prec, m_rows=n * (h + 2) * (h + 2) if k == 3 else n * h * h)
    cw.pack_conv_weight(0, w.to(DEV).contiguous())
    scd = shd = None
    if affine:
        scd = torch.zeros(n, cp, device=DEV); scd[:, :cin] = sc.to(DEV)
        shd = torch.zeros(n, cp, device=DEV); shd[:, :cin] = sh.to(DEV)
    srcs.append(ops.conv_src(ops.to_pnhwc(x.to(DEV), cpad=cp), cp, taps, scale=scd, shift=shd, act=act))
    if skip1x1:
        cw.pack_conv_weight(1, w2.to(DEV).contiguous())
        srcs.append(ops.conv_src(ops.to_pnhwc(xs.to(DEV)), ops.pad_c(skip1x1), ops.TAPS_1X1))
    st = torch.zeros(n, cout, 2, dtype=torch.float64, device=DEV)
    if nchw:
        out = torch.zeros(n, cout, h, h, device=DEV); mode = ops.OUT_NCHW
        rd = r.to(DEV) if res else None
    else:
        out = ops.alloc_pnhwc(n, h, h, cout, DEV); mode = ops.OUT_PNHWC
        rd = ops.to_pnhwc(r.to(DEV), cpad=cout) if res else None
    ops.conv2d_fused(cw, srcs, n, h, h, out, out_mode=mode, bias=b.to(DEV), res=rd,
                     out_scale=(1 / math.sqrt(2) if res else 1.0), stats=st, msub=msub,
                     addvec=(tv.to(DEV) if temb else None), addvec_stride=cout, force_linear=force_linear, splitk_ws=splitk)
    if splitk is not None:
        from ddgan_b200._lib import lib
        assert lib().ddg_conv_last_launch_ksplit() > 1, 'expected a split-K launch'
        assert int(splitk[:4096].view(torch.int32).abs().max()) == 0, 'arrival counters must be back at zero'
    y = out if nchw else ops.from_pnhwc(out, cout)
    err = O.rel_l2(y.cpu(), ref)
    s1 = ref.double().sum(dim=(2, 3)); s2 = (ref.double() ** 2).sum(dim=(2, 3))
    e1 = float((st[:, :, 0].cpu() - s1).norm() / s2.sum().sqrt()); e2 = O.rel_l2(st[:, :, 1].cpu(), s2)
    if not nchw:
        border = out.clone(); border[:, 1:-1, 1:-1, :] = 0
        assert float(border.abs().max()) == 0.0
    return err, e1, e2


@pytest.mark.parametrize('n,cin,cout,h,k', [(2, 32, 128, 8, 1), (2, 32, 128, 8, 3), (4, 64, 128, 16, 3), (8, 128, 256, 32, 3),
                                            (4, 3, 128, 32, 3), (3, 512, 512, 4, 3), (5, 256, 64, 4, 1), (1, 128, 128, 32, 3)])
def test_conv_tc_plain(ops, n, cin, cout, h, k):
    err, e1, e2 = _conv_case(ops, n, cin, cout, h, k)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)


def test_conv_tc_fused_variants(ops):
    err, e1, e2 = _conv_case(ops, 4, 256, 256, 16, 3, affine=True, act=1, res=True, temb=True)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)
    err, _, _ = _conv_case(ops, 4, 128, 3, 32, 3, nchw=True)
    assert err < 2e-5
    err, _, _ = _conv_case(ops, 2, 128, 128, 16, 3, act=2, msub=1)
    assert err < 2e-5
    err, _, _ = _conv_case(ops, 2, 128, 128, 16, 3, act=2, msub=2)
    assert err < 2e-5
    # fused 1x1 skip conv as a second K segment + residual rescale (ResnetBlockBigGANpp_Adagn, layerspp.py:305-310)
    err, e1, e2 = _conv_case(ops, 4, 256, 128, 16, 3, affine=True, act=1, skip1x1=384)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)


@pytest.mark.parametrize('n,cin,cout,h,kw', [(64, 256, 256, 4, dict(affine=True, act=1, res=True, temb=True)),
                                             (64, 256, 256, 4, dict(skip1x1=256, affine=True, act=1)),
                                             (16, 512, 512, 4, dict(act=2)), (8, 256, 256, 8, dict(affine=True, act=1, res=True)),
                                             (2, 128, 128, 16, dict(temb=True)), (64, 128, 128, 4, dict())])
def test_conv_tc_split_k_small_levels(ops, n, cin, cout, h, kw):
    """4x4 / 8x8 levels: a few dozen one-tile CTAs split K between 2 or 4 CTAs per tile and reduce through the workspace; same result
    as the unsplit launch up to fp32 reassociation, statistics included; the workspace is reusable launch after launch."""
    ws = ops.alloc_splitk_ws(DEV)
    base = _conv_case(ops, n, cin, cout, h, 3, **kw)
    for _ in range(2):
        err, e1, e2 = _conv_case(ops, n, cin, cout, h, 3, splitk=ws, **kw)
        assert err < 2e-5 and e1 < 2e-5 and e2 < 2e-5, (err, e1, e2, base)


def test_conv_tc_linear_and_2d_tilings_agree(ops):
    # 16/32-px maps use the 2-D (16 x 8) tiling by default; the 1-D padded-linear tiling must give the same answer
    for (n, cin, cout, h) in [(4, 64, 128, 16), (2, 128, 256, 32), (2, 32, 64, 64)]:
        e2d, _, _ = _conv_case(ops, n, cin, cout, h, 3, affine=True, act=1, res=True)
        e1d, _, _ = _conv_case(ops, n, cin, cout, h, 3, affine=True, act=1, res=True, force_linear=1)
        assert e2d < 2e-5 and e1d < 2e-5, (e2d, e1d)
    for msub in (1, 2):
        err, e1, e2 = _conv_case(ops, 3, 64, 64, 32, 3, msub=msub, skip1x1=64)
        assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5


def test_conv_tc_persistent_variants(ops):
    """More tiles than SMs selects the persistent kernel (several tiles per CTA, TMEM double-buffered epilogue, dedicated epilogue
    warps, producers helping on the last tile).  Shapes are chosen per instantiation; tile counts are not multiples of the SM
    count, so CTAs with different numbers of tiles coexist."""
    # <1,128>: 20 * 8 = 160 tiles of 128 rows, fused 1x1 skip segment
    err, e1, e2 = _conv_case(ops, 20, 64, 128, 32, 3, skip1x1=64)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)
    assert ops.conv_last_launch_info()[:3] == (1, 128, 1), ops.conv_last_launch_info()
    # <2,128>: 56 * 4 = 224 tiles of 256 rows
    err, e1, e2 = _conv_case(ops, 56, 64, 128, 32, 3, affine=True, act=1)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)
    assert ops.conv_last_launch_info()[:3] == (2, 128, 1), ops.conv_last_launch_info()
    # <1,256>: 40 * 8 = 320 tiles, every epilogue feature (AdaGN prologue, temb add, residual, rescale, statistics)
    err, e1, e2 = _conv_case(ops, 40, 128, 256, 32, 3, affine=True, act=1, res=True, temb=True)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)
    assert ops.conv_last_launch_info()[:3] == (1, 256, 1), ops.conv_last_launch_info()
    # <2,64> and <2,16> (NCHW + tanh-less output conv shape)
    err, e1, e2 = _conv_case(ops, 64, 32, 64, 32, 3)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)
    assert ops.conv_last_launch_info()[:3] == (2, 64, 1), ops.conv_last_launch_info()
    err, _, _ = _conv_case(ops, 64, 32, 3, 32, 3, nchw=True)
    assert err < 2e-5
    assert ops.conv_last_launch_info()[:3] == (2, 16, 1), ops.conv_last_launch_info()
    # 1x1 with three N tiles per M tile (the attention QKV projection shape)
    err, e1, e2 = _conv_case(ops, 64, 64, 768, 16, 1, affine=True)
    assert err < 2e-5 and e1 < 5e-5 and e2 < 5e-5, (err, e1, e2)
    assert ops.conv_last_launch_info()[:3] == (1, 256, 1), ops.conv_last_launch_info()
    # single-pass BF16 mode on the persistent path
    err, _, _ = _conv_case(ops, 56, 64, 128, 32, 3, prec=1)
    assert 1e-4 < err < 5e-3, err
    assert ops.conv_last_launch_info()[:3] == (2, 128, 1), ops.conv_last_launch_info()


def test_conv_tc_bf16_mode(ops):
    # BF16 mode: single-pass bf16 operands, fp32 accumulate.  Stated tolerance: 5e-3 relative L2 per op
    # (SURVEY.md section 7 hard part 1 measured 2.3e-3 for bf16 operands).
    err, _, _ = _conv_case(ops, 4, 128, 256, 16, 3, prec=1)
    assert 1e-4 < err < 5e-3, err


def test_batched_gemm_attention_core(ops):
    # S = q k^T * C^-1/2 ; P = softmax(S) ; O = P v   (layerspp.py:115-119) through the batched 1x1 mode
    n, T, c = 3, 64, 128
    q = seeded((n, T, c), 40); k = seeded((n, T, c), 41); v = seeded((n, T, c), 42)
    ref_s = torch.einsum('btc,bsc->bts', q, k) * (c ** -0.5)
    ref_p = torch.softmax(ref_s, dim=-1)
    ref_o = torch.einsum('bts,bsc->btc', ref_p, v)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    wk = ops.ConvWeights(T, [(c, 1)], DEV, batch=n)
    wk.pack_segment(0, kd, c, c, 1, 0, w_batch_stride=T * c)          # B[co = key][ci = channel]
    Tp = ops.pad_c(T)
    s = torch.zeros(n, T, Tp, device=DEV)
    ops.conv2d_fused(wk, [ops.conv_src(qd, c, ops.TAPS_1X1, padded=False)], n, 1, T, s, out_mode=ops.OUT_NHWC, out_c=Tp,
                     out_scale=c ** -0.5, batch_rows=T)
    assert O.rel_l2(s[:, :, :T].cpu(), ref_s) < 2e-5
    p = torch.empty_like(s)
    ops.softmax_rows(s, p, n * T, T, Tp, Tp)
    assert O.rel_l2(p[:, :, :T].cpu(), ref_p) < 2e-5
    wv = ops.ConvWeights(c, [(Tp, 1)], DEV, batch=n)
    wv.pack_segment(0, vd, T, 1, c, 0, w_batch_stride=T * c)          # B[co = channel][ci = key] = v[key][channel]
    o = torch.zeros(n, T, c, device=DEV)
    ops.conv2d_fused(wv, [ops.conv_src(p, Tp, ops.TAPS_1X1, padded=False)], n, 1, T, o, out_mode=ops.OUT_NHWC, batch_rows=T)
    assert O.rel_l2(o.cpu(), ref_o) < 3e-5


@pytest.mark.parametrize('n,prec,tol', [(2, 3, 2e-5), (5, 3, 2e-5), (3, 1, 1e-2)])
def test_fused_attention_core(n, prec, tol):
    """ddg_attention_fwd (QK^T -> softmax -> PV -> NIN_3 + residual, one kernel) against the einsum formulation of
    layerspp.py:115-124 in fp32 on the CPU; 16x16 tokens, 256 channels."""
    from ddgan_b200 import ops
    import math
    H = W = 16; C = 256; T = H * W
    g = torch.Generator().manual_seed(77 + n)
    qkv = torch.randn(n, T, 3 * C, generator=g) * 1.3
    W3 = torch.randn(C, C, generator=g) / math.sqrt(C)
    b3 = torch.randn(C, generator=g) * 0.1
    x = torch.randn(n, C, H, W, generator=g)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    w = torch.softmax(torch.einsum('btc,bsc->bts', q, k) * (C ** -0.5), dim=-1)
    h = torch.einsum('bts,bsc->btc', w, v)
    ref = (x + (h @ W3 + b3).transpose(1, 2).reshape(n, C, H, W)) / math.sqrt(2.0)
    xp = ops.to_pnhwc(x.to(DEV), cpad=C)
    out = ops.alloc_pnhwc(n, H, W, C, DEV)
    st = torch.zeros(n, C, 2, dtype=torch.float64, device=DEV)
    w3 = ops.ConvWeights(C, [(C, 1)], DEV, precision=prec, nt=256)
    w3.pack_nin_weight(0, W3.to(DEV))
    d = ops.attention_desc(qkv.to(DEV), w3, b3.to(DEV), xp, out, st, n, H, W, C, 1.0 / math.sqrt(2.0), precision=prec)
    ops.attention_launch(d)
    y = ops.from_pnhwc(out, C).cpu()
    err = float((y - ref).norm() / ref.norm())
    assert err < tol, err
    if prec == 3:
        assert torch.allclose(st[..., 0].cpu(), ref.double().sum(dim=(2, 3)), rtol=1e-4, atol=1e-3)
        assert torch.allclose(st[..., 1].cpu(), (ref.double() ** 2).sum(dim=(2, 3)), rtol=1e-4, atol=1e-3)
    # border of the PNHWC output stays zero
    assert float(out[:, 0].abs().max()) == 0.0 and float(out[:, :, 0].abs().max()) == 0.0


@pytest.mark.parametrize('n,cin,cskip,cout,h', [(4, 128, 128, 128, 32), (64, 128, 128, 128, 32), (8, 256, 128, 256, 16), (3, 64, 64, 64, 64)])
def test_conv_tma_fed_skip_segments(n, cin, cskip, cout, h):
    """K segments fetched by the TMA engine from pre-split bf16 planes (cp.async.bulk.tensor 5-D boxes) give exactly the result of
    the fp32 producer path: the planes hold the same bf16 hi / lo values the producers would compute.  Layer shape of a resblock
    whose 1x1 skip conv rides as extra K segments: 3x3 over a normalised tensor + two raw 1x1 sources; planes written once by
    ddg_split_planes and once by a producing conv's epilogue (out_planes)."""
    import math
    from ddgan_b200 import ops
    g = torch.Generator().manual_seed(5 + n)
    hx = ops.to_pnhwc(torch.randn(n, cin, h, h, generator=g).to(DEV), cpad=cin)
    xa = ops.to_pnhwc(torch.randn(n, cskip, h, h, generator=g).to(DEV), cpad=cskip)
    xb = ops.to_pnhwc(torch.randn(n, cskip, h, h, generator=g).to(DEV), cpad=cskip)
    sc = (torch.rand(n, cin, generator=g) + 0.5).to(DEV); sh = torch.randn(n, cin, generator=g).to(DEV)
    w3 = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(DEV)
    w1 = (torch.randn(cout, 2 * cskip, generator=g) / math.sqrt(2 * cskip)).to(DEV)
    m_rows = n * (h + 2) * (h + 2)
    cw = ops.ConvWeights(cout, [(cin, 9), (cskip, 1), (cskip, 1)], DEV, m_rows=m_rows)
    cw.pack_conv_weight(0, w3)
    cw.pack_segment(1, w1, cskip, 2 * cskip, 1, 0)
    cw.pack_segment(2, w1, cskip, 2 * cskip, 1, 0, elem_offset=cskip)

    def run(pa, pb):
        out = ops.alloc_pnhwc(n, h, h, cout, DEV)
        srcs = [ops.conv_src(hx, cin, ops.TAPS_3X3, scale=sc, shift=sh, act=ops.ACT_SILU),
                ops.conv_src(xa, cskip, ops.TAPS_1X1, planes=pa), ops.conv_src(xb, cskip, ops.TAPS_1X1, planes=pb)]
        ops.conv2d_fused(cw, srcs, n, h, h, out)
        return out
    ref = run(None, None)
    pa = ops.split_planes(xa, ops.alloc_planes(n, h, h, cskip, 3, DEV), 3)
    # xb's planes come from a conv epilogue: an identity-ish 1x1 conv that reproduces xb exactly is not available, so produce a NEW
    # tensor xb2 with a conv and use it (fp32 + planes from the same epilogue values) in both runs
    wi = (torch.randn(cskip, cskip, generator=g) / math.sqrt(cskip)).to(DEV)
    ci = ops.ConvWeights(cskip, [(cskip, 1)], DEV, m_rows=n * h * h)
    ci.pack_segment(0, wi, cskip, cskip, 1, 0)
    xb2 = ops.alloc_pnhwc(n, h, h, cskip, DEV)
    pb = ops.alloc_planes(n, h, h, cskip, 3, DEV)
    ops.conv2d_fused(ci, [ops.conv_src(xb, cskip, ops.TAPS_1X1)], n, h, h, xb2, out_planes=pb)
    assert torch.equal(ops.split_planes(xb2, ops.alloc_planes(n, h, h, cskip, 3, DEV), 3), pb)     # epilogue planes == split kernel
    xb = xb2
    ref = run(None, None)
    from ddgan_b200._lib import lib
    assert lib().ddg_conv_last_launch_tma() == 0
    got = run(pa, pb)
    assert lib().ddg_conv_last_launch_tma() == 2          # both skip segments were TMA-fed
    assert torch.equal(got, ref), float((got - ref).abs().max())
    # the planes path really ran (the kernel silently falls back to fp32 when the planes are not usable)
    got2 = run(pa, None)
    assert torch.equal(got2, ref)
