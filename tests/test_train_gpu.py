"""Training path on the GPU: gradient parity of the differentiable operators and of the full adversarial step
(ddgan.py:443-518, incl. the R1 double-backward) against the CPU oracle / the reference's own gradients."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ddgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GTOL = 2e-4   # gradients accumulate several BF16x3 GEMMs; still well inside 1e-3


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize('n,cin,cout,h,k', [(2, 32, 64, 8, 3), (4, 64, 128, 16, 3), (3, 128, 32, 8, 1), (2, 256, 256, 4, 3), (8, 128, 128, 32, 3),
                                            (32, 128, 128, 32, 3),    # bench-scale: persistent fwd / dgrad kernels, clustered split-K wgrad
                                            (2, 64, 64, 32, 3), (1, 64, 128, 256, 3),   # 64-channel maps of the 256-px configs (N = 64 wgrad tiles)
                                            (1, 32, 64, 256, 3)])     # narrow input at 256 px: wgrad falls back to one tap row per CTA
def test_conv_grads_and_double_backward(n, cin, cout, h, k):
    from ddgan_b200 import train_graph as TG
    x = seeded((n, cin, h, h), 1); w = seeded((cout, cin, k, k), 2) / math.sqrt(cin * k * k); b = seeded((cout,), 3, 0.1)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=k // 2)
    gy = seeded(tuple(ref.shape), 4)
    # first order + a double-backward scalar: s = sum(g_x^2) with g_x = d(sum(ref*gy))/dx
    gx_ref, = torch.autograd.grad((ref * gy).sum(), xr, create_graph=True)
    (gx_ref ** 2).sum().backward()
    ggw_ref = wr.grad.clone()
    wr.grad = None
    gx1, gw1, gb1 = torch.autograd.grad((F.conv2d(xr, wr, br, padding=k // 2) * gy).sum(), (xr, wr, br))

    xd = x.to(DEV).requires_grad_(True); wd = w.to(DEV).requires_grad_(True); bd = b.to(DEV).requires_grad_(True)
    xp = TG.ToPnhwcFn.apply(xd, cin)
    y = TG.conv3x3(xp, wd, bd, n, h, h) if k == 3 else TG.conv1x1(xp, wd, bd, n, h, h)
    yo = TG.FromPnhwcFn.apply(y, cout)
    assert O.rel_l2(yo.detach().cpu(), ref.detach()) < 2e-5
    gx, gw, gb = torch.autograd.grad((yo * gy.to(DEV)).sum(), (xd, wd, bd), create_graph=True)
    assert O.rel_l2(gx.detach().cpu(), gx1) < 5e-5
    assert O.rel_l2(gw.detach().cpu(), gw1) < 5e-5
    assert O.rel_l2(gb.detach().cpu(), gb1) < 5e-5
    ggw, = torch.autograd.grad((gx ** 2).sum(), wd)
    assert O.rel_l2(ggw.cpu(), ggw_ref) < GTOL


def test_conv_residual_and_rescale_in_epilogue():
    """y = s * (conv(x) + b + res): the residual and the 1/sqrt(2) rescale ride in the conv epilogue; s is the dgrad launch's
    epilogue scale and the wgrad reduction's gain (layerspp.py:305-310)."""
    from ddgan_b200 import train_graph as TG
    n, cin, cout, h, sc = 4, 64, 128, 16, 1.0 / math.sqrt(2.0)
    x = seeded((n, cin, h, h), 61); w = seeded((cout, cin, 3, 3), 62) / math.sqrt(cin * 9); b = seeded((cout,), 63, 0.1)
    r = seeded((n, cout, h, h), 64)
    xr, wr, br, rr = [t.clone().requires_grad_(True) for t in (x, w, b, r)]
    ref = (F.conv2d(xr, wr, br, padding=1) + rr) * sc
    gy = seeded(tuple(ref.shape), 65)
    g_ref = torch.autograd.grad((ref * gy).sum(), (xr, wr, br, rr))
    xd, wd, bd, rd = [t.to(DEV).requires_grad_(True) for t in (x, w, b, r)]
    y = TG.conv3x3(TG.ToPnhwcFn.apply(xd, cin), wd, bd, n, h, h, res=TG.ToPnhwcFn.apply(rd, cout), out_scale=sc)
    yo = TG.FromPnhwcFn.apply(y, cout)
    assert O.rel_l2(yo.detach().cpu(), ref.detach()) < 2e-5
    g = torch.autograd.grad((yo * gy.to(DEV)).sum(), (xd, wd, bd, rd))
    for a, e in zip(g, g_ref):
        assert O.rel_l2(a.cpu(), e) < 5e-5


def test_groupnorm_fir_linear_grads():
    from ddgan_b200 import train_graph as TG
    from ddgan_b200 import ops
    n, c, h = 3, 64, 8
    x = seeded((n, c, h, h), 5) * 1.5 + 0.3
    gamma = 1 + seeded((n, c), 6) * 0.2; beta = seeded((n, c), 7) * 0.2
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.silu(gr[:, :, None, None] * O.group_norm(xr, O.num_groups(c)) + br[:, :, None, None])
    ref = O.downsample_2d(O.upsample_2d(ref))
    gy = seeded(tuple(ref.shape), 8)
    g_ref = torch.autograd.grad((ref * gy).sum(), (xr, gr, br))
    xd, gd, bd = x.to(DEV).requires_grad_(True), gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y = TG.group_norm_act(TG.ToPnhwcFn.apply(xd, c), h, h, O.num_groups(c), gd, bd, ops.ACT_SILU)
    y = TG.fir_down(TG.fir_up(y))
    yo = TG.FromPnhwcFn.apply(y, c)
    assert O.rel_l2(yo.detach().cpu(), ref.detach()) < 1e-5
    g = torch.autograd.grad((yo * gy.to(DEV)).sum(), (xd, gd, bd))
    for a, b_ in zip(g, g_ref):
        assert O.rel_l2(a.cpu(), b_) < 5e-5
    # the fused node the generator's training path uses (GnActFn: 3 + 3 launches), AdaGN form: [gamma | beta] are columns of a
    # wider projection tensor; and the shared-affine form of the attention / output norms
    xd, gd, bd = x.to(DEV).requires_grad_(True), gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    pad = torch.zeros(n, 8, device=DEV)
    style = torch.cat([pad, gd, bd, pad], 1)
    y = TG.gn_act(TG.ToPnhwcFn.apply(xd, c), h, h, O.num_groups(c), ops.ACT_SILU, style=style, off=8)
    yo = TG.FromPnhwcFn.apply(TG.fir_down(TG.fir_up(y)), c)
    assert O.rel_l2(yo.detach().cpu(), ref.detach()) < 1e-5
    g = torch.autograd.grad((yo * gy.to(DEV)).sum(), (xd, gd, bd))
    for a, b_ in zip(g, g_ref):
        assert O.rel_l2(a.cpu(), b_) < 5e-5
    ga, be = 1 + seeded((c,), 16) * 0.2, seeded((c,), 17) * 0.2
    xr, gr, br = x.clone().requires_grad_(True), ga.clone().requires_grad_(True), be.clone().requires_grad_(True)
    ref2 = O.group_norm(xr, O.num_groups(c), weight=gr, bias=br)
    gy2 = seeded(tuple(ref2.shape), 18)
    g_ref2 = torch.autograd.grad((ref2 * gy2).sum(), (xr, gr, br))
    xd, gd, bd = x.to(DEV).requires_grad_(True), ga.to(DEV).requires_grad_(True), be.to(DEV).requires_grad_(True)
    yo = TG.FromPnhwcFn.apply(TG.gn_act(TG.ToPnhwcFn.apply(xd, c), h, h, O.num_groups(c), ops.ACT_NONE, gamma=gd, beta=bd), c)
    assert O.rel_l2(yo.detach().cpu(), ref2.detach()) < 1e-5
    g = torch.autograd.grad((yo * gy2.to(DEV)).sum(), (xd, gd, bd))
    for a, b_ in zip(g, g_ref2):
        assert O.rel_l2(a.cpu(), b_) < 5e-5
    # linear
    xl = seeded((64, 100), 9); W = seeded((256, 100), 10, 0.1); b = seeded((256,), 11)
    xr, Wr, br = xl.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    gy = seeded((64, 256), 12)
    g_ref = torch.autograd.grad((F.linear(xr, Wr, br) * gy).sum(), (xr, Wr, br))
    xd, Wd, bd = xl.to(DEV).requires_grad_(True), W.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    g = torch.autograd.grad((TG.LinearFn.apply(xd, Wd, bd) * gy.to(DEV)).sum(), (xd, Wd, bd))
    for a, b_ in zip(g, g_ref):
        assert O.rel_l2(a.cpu(), b_) < 1e-5
    # wide linear (the batched AdaGN style projection shape): forward and both gradient GEMMs run on the tcgen05 kernels
    xl = seeded((64, 256), 13); W = seeded((4096, 256), 14, 0.1); b = seeded((4096,), 15)
    assert TG._tc_linear(64, 256, 4096)
    xr, Wr, br = xl.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    gy = seeded((64, 4096), 16)
    y_ref = F.linear(xr, Wr, br)
    g_ref = torch.autograd.grad((y_ref * gy).sum(), (xr, Wr, br))
    xd, Wd, bd = xl.to(DEV).requires_grad_(True), W.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    yd = TG.LinearFn.apply(xd, Wd, bd)
    assert O.rel_l2(yd.detach().cpu(), y_ref.detach()) < 1e-5
    g = torch.autograd.grad((yd * gy.to(DEV)).sum(), (xd, Wd, bd))
    for a, b_ in zip(g, g_ref):
        assert O.rel_l2(a.cpu(), b_) < 1e-5


def _nets():
    from ddgan_b200.modules import NCSNpp, Discriminator_small
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
    netG = NCSNpp(cfg).to(DEV)
    netG.load_state_dict(O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21))
    netD = Discriminator_small(nc=6, ngf=16, t_emb_dim=32).to(DEV)
    netD.load_state_dict(O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22))
    return cfg, netG, netD


def test_adversarial_step_gradients_vs_reference(golden):
    """The D step (real + R1 + fake, three backward calls accumulating into .grad) and the G step of ddgan.py:443-508 with the
    reference's own gradients (tests/golden/make_golden.py) as the expected values."""
    from ddgan_b200 import diffusion
    g = golden['train_tiny']
    cfg, netG, netD = _nets()
    coeff = diffusion.DiffusionCoefficients(cfg, DEV); pc = diffusion.PosteriorCoefficients(cfg, DEV)
    real = torch.tanh(seeded((4, 3, 32, 32), 300)).to(DEV); t = torch.tensor([0, 1, 2, 3], device=DEV)
    n_xtp1, n_xt, n_post = (seeded((4, 3, 32, 32), s).to(DEV) for s in (301, 302, 303))
    z = seeded((4, cfg.nz), 304).to(DEV)
    x_t, x_tp1 = diffusion.q_sample_pairs(coeff, real, t, noise_xt=n_xt, noise_xtp1=n_xtp1)
    x_t.requires_grad = True
    netD.zero_grad(); netG.zero_grad()
    D_real = netD(x_t, t, x_tp1.detach()).view(-1)
    errD_real = F.softplus(-D_real).mean()
    errD_real.backward(retain_graph=True)
    grad_real = torch.autograd.grad(outputs=D_real.sum(), inputs=x_t, create_graph=True)[0]
    gp = g['r1_gamma'] / 2 * (grad_real.view(grad_real.size(0), -1).norm(2, dim=1) ** 2).mean()
    gp.backward()
    x0p = netG(x_tp1.detach(), t, z)
    x_pos = diffusion.sample_posterior(pc, x0p, x_tp1, t, noise=n_post)
    output = netD(x_pos, t, x_tp1.detach()).view(-1)
    errD_fake = F.softplus(output).mean()
    errD_fake.backward()
    assert abs(float(errD_real) - float(g['errD_real'])) < 1e-4 * max(1, abs(float(g['errD_real'])))
    assert abs(float(gp) - float(g['gp'])) < 5e-4 * abs(float(g['gp']))
    assert abs(float(errD_fake) - float(g['errD_fake'])) < 1e-4 * max(1, abs(float(g['errD_fake'])))
    assert O.rel_l2(x_pos.detach().cpu(), g['x_pos']) < 1e-4
    gd = dict(netD.named_parameters())
    worst = max(O.rel_l2(gd[k].grad.cpu(), v) for k, v in g['gradD'].items())
    assert worst < 5e-4, worst
    gg = dict(netG.named_parameters())
    worst = max(O.rel_l2(gg[k].grad.cpu(), v) for k, v in g['gradG_in_dstep'].items())
    assert worst < 5e-4, worst
    # G step
    netG.zero_grad()
    for p in netD.parameters():
        p.requires_grad = False
    x0p = netG(x_tp1.detach(), t, z)
    x_pos = diffusion.sample_posterior(pc, x0p, x_tp1, t, noise=n_post)
    errG = F.softplus(-netD(x_pos, t, x_tp1.detach()).view(-1)).mean()
    errG.backward()
    assert abs(float(errG) - float(g['errG'])) < 1e-4 * max(1, abs(float(g['errG'])))
    worst = max(O.rel_l2(gg[k].grad.cpu(), v) for k, v in g['gradG'].items())
    assert worst < 5e-4, worst
    tot = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in netG.parameters()))
    assert abs(float(tot) - float(g['gradG_norm'])) < 5e-4 * float(g['gradG_norm'])


def test_skipping_the_discarded_generator_backward_is_exact():
    """Losses and both parameter updates agree with and without the generator backward that ddgan.py:489 throws away (same
    injected randomness), up to the summation order of atomic reductions."""
    import copy
    from ddgan_b200.train import Trainer
    cfg, netG, netD = _nets()
    for k, v in dict(lr_g=1.6e-4, lr_d=1.25e-4, beta1=0.5, beta2=0.9, r1_gamma=0.02, lazy_reg=1, grad_clip_norm=1.0,
                     ema_decay=0.999, use_ema=True).items():
        setattr(cfg, k, v)
    netG2, netD2 = copy.deepcopy(netG), copy.deepcopy(netD)
    real = torch.tanh(seeded((4, 3, 32, 32), 300)).to(DEV)
    nz = {}
    for sfx, base in (('_d', 400), ('_g', 500)):
        nz['t' + sfx] = torch.tensor([0, 1, 2, 3], device=DEV)
        nz['n_xtp1' + sfx] = seeded((4, 3, 32, 32), base + 1).to(DEV); nz['n_xt' + sfx] = seeded((4, 3, 32, 32), base + 2).to(DEV)
        nz['z' + sfx] = seeded((4, cfg.nz), base + 3).to(DEV); nz['n_post' + sfx] = seeded((4, 3, 32, 32), base + 4).to(DEV)
    a = Trainer(cfg, netG, netD, DEV, skip_discarded_g_backward=True)
    b = Trainer(cfg, netG2, netD2, DEV, skip_discarded_g_backward=False)
    ea = a.step(real, 0, noise=nz)
    eb = b.step(real, 0, noise=nz)
    assert abs(float(ea[0]) - float(eb[0])) < 1e-5 and abs(float(ea[1]) - float(eb[1])) < 1e-5
    # Both trainers run the same schedule; the only numerical difference is that the skipped pass goes through the fused
    # inference engine and the faithful one through the differentiable path (fake images equal to ~1e-5).  LeakyReLU gating
    # makes D's gradient discontinuous in its input, so that 1e-5 becomes ~1e-3 in gradient space (a flipped gate changes a
    # whole path), and the first Adam step (lr * g / (|g| + eps)) turns it into +-lr flips of the few ~0-gradient elements.
    assert O.rel_l2(a.optD.flat_g.cpu(), b.optD.flat_g.cpu()) < 5e-3
    assert O.rel_l2(a.optG.flat_g.cpu(), b.optG.flat_g.cpu()) < 5e-3
    for (n1, p1), (n2, p2) in zip(list(netD.named_parameters()) + list(netG.named_parameters()),
                                  list(netD2.named_parameters()) + list(netG2.named_parameters())):
        assert O.rel_l2(p1.detach().cpu(), p2.detach().cpu()) < 1e-3, n1


def test_flat_adam_matches_torch_clip_adam_ema():
    """ddg_grad_norm_sq + ddg_adam_ema_step vs clip_grad_norm_ + torch.optim.Adam + the EMA update of ema.py:45-55."""
    from ddgan_b200.train import FlatAdam
    torch.manual_seed(3)
    net_a = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 11)).to(DEV)
    import copy
    net_b = copy.deepcopy(net_a)
    fa = FlatAdam(net_a, 1.6e-4, (0.5, 0.9), max_norm=1.0, ema_decay=0.999)
    opt = torch.optim.Adam(net_b.parameters(), lr=1.6e-4, betas=(0.5, 0.9))
    ema = [p.detach().clone() for p in net_b.parameters()]
    for it in range(4):
        x = torch.randn(16, 37, device=DEV) * (10.0 if it % 2 else 0.01)   # exercise both clipped and unclipped steps
        fa.zero_grad(); opt.zero_grad()
        (net_a(x) ** 2).sum().backward()
        (net_b(x) ** 2).sum().backward()
        fa.step()
        torch.nn.utils.clip_grad_norm_(net_b.parameters(), 1.0)
        opt.step()
        for e, p in zip(ema, net_b.parameters()):
            e.mul_(0.999).add_(p.detach(), alpha=0.001)
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert O.rel_l2(pa.detach().cpu(), pb.detach().cpu()) < 1e-6
    sd = fa.ema_state_dict(net_a)
    for (n, _), e in zip(net_a.named_parameters(), ema):
        assert O.rel_l2(sd[n].cpu(), e.cpu()) < 1e-6


@pytest.mark.parametrize('n,t,c', [(3, 256, 256), (4, 16, 64), (2, 64, 128)])
def test_attention_core_function_grads(n, t, c):
    """AttnCoreFn (training path of layerspp.py:115-119 on this library's GEMM / softmax kernels) against torch autograd."""
    from ddgan_b200 import train_graph as TG
    q, k, v = seeded((n, t, c), 31), seeded((n, t, c), 32), seeded((n, t, c), 33)
    gy = seeded((n, t, c), 34)
    qr, kr, vr = [x.clone().requires_grad_(True) for x in (q, k, v)]
    ref = torch.bmm(torch.softmax(torch.bmm(qr, kr.transpose(1, 2)) * (c ** -0.5), dim=-1), vr)
    g_ref = torch.autograd.grad((ref * gy).sum(), (qr, kr, vr))
    qd, kd, vd = [x.to(DEV).requires_grad_(True) for x in (q, k, v)]
    o = TG.AttnCoreFn.apply(qd, kd, vd, 3)
    assert O.rel_l2(o.detach().cpu(), ref.detach()) < 2e-5
    g = torch.autograd.grad((o * gy.to(DEV)).sum(), (qd, kd, vd))
    for a, b_ in zip(g, g_ref):
        assert O.rel_l2(a.cpu(), b_) < 5e-5
