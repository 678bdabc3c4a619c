"""Parity at the BASELINE.json configurations (not shrunken ones): Discriminator_small at ngf 64 / t_emb 256 / batch 64,
Discriminator_large against the reference's own golden output, a T = 4 sampled batch of 64 CIFAR images, one CIFAR-config
adversarial step (losses + every gradient tensor, R1 included) and the CUDA-graph train step against the eager one.

Expected values come from the CPU oracle (pinned to the reference by tests/test_oracle_golden.py) or from the committed
reference golden directly.  FP32-mode gate: 1e-4 relative L2 on forward quantities; gradient tolerances are stated per test."""
import copy

import pytest
import torch
import torch.nn.functional as F

from oracle import ddgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-4


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def test_discriminator_small_baseline_config_b64():
    """BASELINE configs[1] discriminator: ngf 64, t_emb_dim 256, 64 samples (16 stddev sets of 4)."""
    from ddgan_b200.engine import DiscriminatorEngine
    B = 64
    sd = O.randomize_params(O.discriminator_param_shapes(6, 64, 256), seed=41)
    x = seeded((B, 3, 32, 32), 420); xt = seeded((B, 3, 32, 32), 421); t = torch.arange(B) % 4
    stages = []
    with torch.no_grad():
        ref = O.discriminator_forward(sd, x, t, xt, 256, stages=stages)
    eng = DiscriminatorEngine(6, 64, 256, 32, B, large=False, device=DEV)
    eng.load_state_dict(sd)
    eng.capture()
    y = eng.forward(x.to(DEV), t.to(DEV), xt.to(DEV)).cpu()
    # every tensor-valued stage meets the 1e-4 gate (measured: 4e-6 after start_conv growing to 2.8e-5 after final_conv) ...
    from ddgan_b200 import ops
    for i, act in enumerate(eng.stage_acts):
        assert O.rel_l2(ops.from_pnhwc(act.buf, stages[i].shape[1]).cpu(), stages[i]) < TOL, i
    feat = F.leaky_relu(ops.from_pnhwc(eng.final_feat.buf, stages[5].shape[1]).cpu(), 0.2)
    assert O.rel_l2(feat, stages[5]) < TOL
    pooled = eng.pooled.cpu()
    assert O.rel_l2(pooled, stages[6]) < TOL
    # ... and the logit is the 512-term dot product end_linear(pooled), whose terms cancel (|w|.|pooled| is ~5x |w.pooled| here):
    # the backward-stable bound for a 1e-4-accurate pooled vector is 1e-4 * (|w| . |pooled| + |b|) per sample.
    w, b = sd['end_linear.weight'], sd['end_linear.bias']
    bound = TOL * (stages[6].abs() @ w.abs().t() + b.abs())
    assert bool(((y - ref).abs() <= bound).all())
    assert O.rel_l2(y, ref) < 5e-4


def test_discriminator_large_vs_reference_golden(golden):
    """The reference's own Discriminator_large(nc=6, ngf=8, t_emb_dim=32) output on 256-px inputs (make_golden.py): 16-channel
    maps, i.e. the narrow-map path (buffers padded to 32 channels)."""
    from ddgan_b200.engine import DiscriminatorEngine
    sd = O.randomize_params(golden['dlarge_shapes'], seed=10)
    x = seeded((4, 3, 256, 256), 230); xt = seeded((4, 3, 256, 256), 231); t = torch.tensor([0, 1, 2, 3])
    eng = DiscriminatorEngine(6, 8, 32, 256, 4, large=True, device=DEV)
    eng.load_state_dict(sd)
    y = eng.forward(x.to(DEV), t.to(DEV), xt.to(DEV))
    assert O.rel_l2(y.cpu(), golden['dlarge_out']) < TOL
    # and through the drop-in module (state_dict names of the reference, strict load)
    from ddgan_b200.modules import Discriminator_large
    net = Discriminator_large(nc=6, ngf=8, t_emb_dim=32).to(DEV).eval()
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        y = net(x.to(DEV), t.to(DEV), xt.to(DEV))
    assert O.rel_l2(y.cpu(), golden['dlarge_out']) < TOL


def test_sampled_images_baseline_config_b64():
    """BASELINE configs[0]: 64 CIFAR images, T = 4, injected noise (the reference's draw order: z then posterior noise per
    step); the whole loop runs as the CUDA graph bench.py times."""
    from ddgan_b200.engine import GeneratorEngine
    from ddgan_b200 import diffusion
    cfg = O.cifar10_config()
    B = 64
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=43)
    draws = []
    g = torch.Generator().manual_seed(44)

    def noise_fn(shape):
        d = torch.randn(*shape, generator=g)
        draws.append(d)
        return d
    x_init = seeded((B, 3, 32, 32), 45)
    ref = O.sample_from_model(O.posterior_coefficients(cfg), lambda x, t, z: O.ncsnpp_forward(sd, cfg, x, t, z), cfg.num_timesteps,
                              x_init, cfg.nz, noise_fn)
    eng = GeneratorEngine(cfg, B, DEV)
    eng.load_state_dict(sd)
    smp = diffusion.GraphSampler(eng, cfg)
    smp.capture()
    for k in range(cfg.num_timesteps):
        smp.z_noise[k].copy_(draws[2 * k]); smp.p_noise[k].copy_(draws[2 * k + 1])
    y = smp.sample(x_init.to(DEV), fresh_noise=False)
    err = O.rel_l2(y.cpu(), ref)
    per_image = ((y.cpu() - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max()
    assert err < TOL and float(per_image) < TOL, (err, float(per_image))


def _cifar_nets(seed_g=51, seed_d=52):
    from ddgan_b200.modules import NCSNpp, Discriminator_small
    cfg = O.cifar10_config()
    sd_g = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=seed_g)
    sd_d = O.randomize_params(O.discriminator_param_shapes(6, 64, 256), seed=seed_d)
    netG = NCSNpp(cfg).to(DEV)
    netG.load_state_dict(sd_g, strict=True)
    netD = Discriminator_small(nc=6, ngf=64, t_emb_dim=256).to(DEV)
    netD.load_state_dict(sd_d, strict=True)
    return cfg, netG, netD, sd_g, sd_d


def test_adversarial_step_baseline_config_gradients_vs_oracle():
    """One iteration of ddgan.py:443-508 at the CIFAR configuration (NCSN++ ch 128 / 1-2-2-2, Discriminator_small ngf 64),
    batch 4, R1 on: losses, and every gradient tensor of both networks, against CPU autograd through the oracle.

    Tolerances (relative L2): losses 1e-4 (R1 penalty 5e-4: it is a sum of squared gradients, i.e. twice the relative error of
    a gradient); whole-network gradient 5e-4; individual tensors 2e-3.  The per-tensor bound is looser than the forward gate
    because D's LeakyReLU gates make its gradient discontinuous in the activations: a 1e-5 forward difference flips a few
    gates, and each flip changes a whole back-propagated path, which shows up in small tensors (biases) first.  Tensors
    whose true gradient is zero (the attention key bias NIN_1.b: softmax is invariant to a per-row shift of the logits) hold
    only rounding noise on both sides; they are bounded in absolute terms against the network's gradient scale instead."""
    from ddgan_b200 import diffusion
    cfg, netG, netD, sd_g, sd_d = _cifar_nets()
    B, gamma = 4, 0.02
    real = torch.tanh(seeded((B, 3, 32, 32), 500)); t = torch.tensor([0, 1, 2, 3])
    n_xt, n_xtp1, n_post = seeded((B, 3, 32, 32), 501), seeded((B, 3, 32, 32), 502), seeded((B, 3, 32, 32), 503)
    z = seeded((B, cfg.nz), 504)
    # ---- oracle (CPU autograd) ----
    pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    pd = {k: v.clone().requires_grad_(True) for k, v in sd_d.items()}
    er, gp, ef = O.d_step_losses(pg, pd, cfg, real, t, (n_xt, n_xtp1, n_post), z, gamma, do_r1=True)
    (er + gp + ef).backward()
    ref_gD = {k: v.grad.clone() for k, v in pd.items()}
    for v in list(pg.values()) + list(pd.values()):
        v.grad = None
    eg = O.g_step_loss(pg, {k: v.detach() for k, v in pd.items()}, cfg, real, t, (n_xt, n_xtp1, n_post), z)
    eg.backward()
    ref_gG = {k: v.grad.clone() for k, v in pg.items()}
    # ---- this framework: the statements of ddgan.py:449-503 on the drop-in modules ----
    coeff = diffusion.DiffusionCoefficients(cfg, DEV); pc = diffusion.PosteriorCoefficients(cfg, DEV)
    td = t.to(DEV)
    x_t, x_tp1 = diffusion.q_sample_pairs(coeff, real.to(DEV), td, noise_xt=n_xt.to(DEV), noise_xtp1=n_xtp1.to(DEV))
    x_t.requires_grad = True
    netD.zero_grad(); netG.zero_grad()
    D_real = netD(x_t, td, x_tp1.detach()).view(-1)
    errD_real = F.softplus(-D_real).mean()
    errD_real.backward(retain_graph=True)
    grad_real = torch.autograd.grad(outputs=D_real.sum(), inputs=x_t, create_graph=True)[0]
    grad_penalty = gamma / 2 * (grad_real.view(B, -1).norm(2, dim=1) ** 2).mean()
    grad_penalty.backward()
    with torch.no_grad():
        x0p = netG(x_tp1.detach(), td, z.to(DEV))
        x_pos = diffusion.sample_posterior(pc, x0p, x_tp1.detach(), td, noise=n_post.to(DEV))
    errD_fake = F.softplus(netD(x_pos, td, x_tp1.detach()).view(-1)).mean()
    errD_fake.backward()
    assert abs(float(errD_real) - float(er)) < 1e-4 * max(1.0, abs(float(er)))
    assert abs(float(errD_fake) - float(ef)) < 1e-4 * max(1.0, abs(float(ef)))
    assert abs(float(grad_penalty) - float(gp)) < 5e-4 * abs(float(gp))
    gd = {k: p.grad.detach().cpu() for k, p in netD.named_parameters()}
    flat = lambda d, keys: torch.cat([d[k].flatten().double() for k in keys])
    def check(got, ref):
        keys = list(ref)
        total = float(flat(ref, keys).norm())
        assert O.rel_l2(flat(got, keys), flat(ref, keys)) < 5e-4
        live = [k for k in keys if float(ref[k].double().norm()) > 1e-6 * total]
        dead = [k for k in keys if k not in live]
        worst = max((O.rel_l2(got[k], ref[k]), k) for k in live)
        assert worst[0] < 2e-3, worst
        for k in dead:
            assert float(got[k].double().norm()) < 1e-5 * total, k
        return len(live), len(dead)
    check(gd, ref_gD)
    # G step
    for p in netD.parameters():
        p.requires_grad = False
    netG.zero_grad()
    x0p = netG(x_tp1.detach(), td, z.to(DEV))
    x_pos = diffusion.sample_posterior(pc, x0p, x_tp1, td, noise=n_post.to(DEV))
    errG = F.softplus(-netD(x_pos, td, x_tp1.detach()).view(-1)).mean()
    errG.backward()
    assert abs(float(errG) - float(eg)) < 1e-4 * max(1.0, abs(float(eg)))
    gg = {k: p.grad.detach().cpu() for k, p in netG.named_parameters()}
    n_live, n_dead = check(gg, ref_gG)
    assert n_dead <= 8, n_dead       # the four NIN_1.b tensors (and nothing substantial) are the zero-gradient ones


def _train_cfg(cfg, lazy_reg=2):
    for k, v in dict(lr_g=1.6e-4, lr_d=1.25e-4, beta1_g=0.5, beta2_g=0.9, beta1_d=0.5, beta2_d=0.9, weight_decay_G=0.0,
                     weight_decay_D=0.0, r1_gamma=0.02, lazy_reg=lazy_reg, grad_clip_norm=1.0, ema_decay=0.999,
                     use_ema=True).items():
        setattr(cfg, k, v)
    return cfg


def _tiny_nets():
    from ddgan_b200.modules import NCSNpp, Discriminator_small
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
    netG = NCSNpp(cfg).to(DEV)
    netG.load_state_dict(O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21))
    netD = Discriminator_small(nc=6, ngf=16, t_emb_dim=32).to(DEV)
    netD.load_state_dict(O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22))
    return cfg, netG, netD


def _noise(cfg, B, S, base):
    nz = {}
    for sfx, b in (('_d', base), ('_g', base + 50)):
        nz['t' + sfx] = (torch.arange(B) % cfg.num_timesteps).to(DEV)
        nz['n_xtp1' + sfx] = seeded((B, 3, S, S), b + 1).to(DEV); nz['n_xt' + sfx] = seeded((B, 3, S, S), b + 2).to(DEV)
        nz['z' + sfx] = seeded((B, cfg.nz), b + 3).to(DEV); nz['n_post' + sfx] = seeded((B, 3, S, S), b + 4).to(DEV)
    return nz


def _sync_state(a, b):
    for oa, ob in ((a.optD, b.optD), (a.optG, b.optG)):
        for ta, tb in ((oa.flat_p, ob.flat_p), (oa.m, ob.m), (oa.v, ob.v), (oa.state, ob.state)):
            tb.copy_(ta)
    b.optG.ema.copy_(a.optG.ema)


def _update_rel(pa, pb, p_before):
    """|p_b - p_a| relative to the size of the update itself (Adam's first steps are sign-like: lr * g / (|g| + eps), so
    elements with |g| ~ eps flip between +-lr on 1e-7 gradient noise; relative to |p| that is invisible, relative to the update it
    is the honest measure)."""
    return float((pb - pa).double().norm() / (pa - p_before).double().norm().clamp_min(1e-30))


@pytest.mark.parametrize('which', ['tiny', 'cifar'])
def test_graphed_train_step_equals_eager_step(which):
    """The benchmarked path (Trainer.step_graphed: two captured CUDA graphs, with and without the lazy R1 term) against the
    eager Trainer.step on identical weights, data and injected randomness; every iteration starts from identical state.

    What can differ between two runs of the SAME step is the summation order of fp32 atomics (split-K weight gradients):
    ~1e-7 per element.  Through the D update that noise reaches the G step of the same iteration, where it occasionally flips
    a LeakyReLU gate of D and changes G's gradient discretely by 1e-4 .. 6e-4 in some tensors -- two eager runs differ from
    each other in exactly this way (tools/train_repeat.py, profiles/r2_train_repeatability.txt).  To test the graph and not
    that chaos, iterations alternate which network's learning rate is zero (the rate lives in device memory, so the same
    graphs are replayed):
      lr_d = 0 (R1 and plain variant): the G step sees an unchanged D -> losses, BOTH gradient arenas, the G update and the
               EMA must agree to 1e-5 (update-relative 1e-2, see _update_rel);
      lr_g = 0: D gradients, D update to the same bounds.
    Also checks that capture() itself leaves weights, Adam state and EMA untouched."""
    from ddgan_b200.train import Trainer
    if which == 'tiny':
        cfg, netG, netD = _tiny_nets()
        B = 4
    else:
        cfg, netG, netD, _, _ = _cifar_nets()
        B = 8
    cfg = _train_cfg(cfg, lazy_reg=2)
    netG2, netD2 = copy.deepcopy(netG), copy.deepcopy(netD)
    real = torch.tanh(seeded((B, 3, 32, 32), 300)).to(DEV)
    a = Trainer(cfg, netG, netD, DEV)
    b = Trainer(cfg, netG2, netD2, DEV)
    p0 = b.optG.flat_p.clone(); d0 = b.optD.flat_p.clone()
    b.capture((B, 3, 32, 32), warmup=3)
    assert torch.equal(b.optG.flat_p, p0) and torch.equal(b.optD.flat_p, d0)
    assert float(b.optG.state[0]) == 0.0 and float(b.optD.state[0]) == 0.0
    assert float(b.optG.m.abs().max()) == 0.0 and float(b.optD.v.abs().max()) == 0.0
    assert torch.equal(b.optG.ema, p0)
    for it, frozen in enumerate(['d', 'd', 'g', 'g']):            # it = 0, 2: R1 graph; it = 1, 3: plain graph
        for tr in (a, b):
            tr.optD.set_lr(0.0 if frozen == 'd' else cfg.lr_d)
            tr.optG.set_lr(0.0 if frozen == 'g' else cfg.lr_g)
        pg, pd_ = a.optG.flat_p.clone(), a.optD.flat_p.clone()
        nz = _noise(cfg, B, 32, 600 + 10 * it)
        ea = a.step(real, it, noise=nz)
        eb = b.step_graphed(real, it, noise=nz)
        assert abs(float(ea[0]) - float(eb[0])) < 1e-5 * max(1.0, abs(float(ea[0]))), (it, float(ea[0]), float(eb[0]))
        assert abs(float(ea[1]) - float(eb[1])) < 1e-5 * max(1.0, abs(float(ea[1]))), (it, float(ea[1]), float(eb[1]))
        assert O.rel_l2(b.optD.flat_g.cpu(), a.optD.flat_g.cpu()) < 1e-5, it
        if frozen == 'd':
            assert O.rel_l2(b.optG.flat_g.cpu(), a.optG.flat_g.cpu()) < 1e-5, it
            assert torch.equal(a.optD.flat_p, pd_) and torch.equal(b.optD.flat_p, pd_)
            assert O.rel_l2(b.optG.flat_p.cpu(), a.optG.flat_p.cpu()) < 1e-5, it
            assert _update_rel(a.optG.flat_p, b.optG.flat_p, pg) < 1e-2, it
            assert O.rel_l2(b.optG.ema.cpu(), a.optG.ema.cpu()) < 1e-5, it
        else:
            assert torch.equal(a.optG.flat_p, pg) and torch.equal(b.optG.flat_p, pg)
            assert O.rel_l2(b.optD.flat_p.cpu(), a.optD.flat_p.cpu()) < 1e-5, it
            assert _update_rel(a.optD.flat_p, b.optD.flat_p, pd_) < 1e-2, it
        _sync_state(a, b)
    assert float(a.optG.state[0]) == 4.0 and float(b.optG.state[0]) == 4.0


def test_engine_is_not_stale_after_graph_replays_and_ema_swap():
    """CUDA-graph replays and EMA swaps change the parameters behind autograd's back; a later no_grad call at
    another batch size (a different cached engine) must see the current weights."""
    from ddgan_b200.train import Trainer
    from ddgan_b200.engine import GeneratorEngine
    cfg, netG, netD = _tiny_nets()
    cfg = _train_cfg(cfg)
    B = 4
    real = torch.tanh(seeded((B, 3, 32, 32), 300)).to(DEV)
    x = seeded((2, 3, 32, 32), 310).to(DEV); t = torch.tensor([1, 3], device=DEV); z = seeded((2, cfg.nz), 311).to(DEV)
    tr = Trainer(cfg, netG, netD, DEV)
    netG.eval()
    with torch.no_grad():
        y0 = netG(x, t, z).clone()          # builds and packs the batch-2 engine from the initial weights
    netG.train()
    tr.capture((B, 3, 32, 32), warmup=3)
    for it in range(4):
        tr.step_graphed(real, it)
    netG.eval()

    def fresh():
        eng = GeneratorEngine(cfg, 2, DEV)
        eng.load_state_dict({k: p.detach().clone() for k, p in netG.named_parameters()})
        return eng.forward(x, t, z).clone()
    with torch.no_grad():
        y1 = netG(x, t, z).clone()
    assert O.rel_l2(y1.cpu(), fresh().cpu()) < 1e-6
    assert O.rel_l2(y1.cpu(), y0.cpu()) > 1e-6      # four Adam steps did move the output
    tr.swap_parameters_with_ema(store_params_in_ema=True)
    with torch.no_grad():
        y2 = netG(x, t, z).clone()
    assert O.rel_l2(y2.cpu(), fresh().cpu()) < 1e-6
    assert O.rel_l2(y2.cpu(), y1.cpu()) > 1e-7
    tr.swap_parameters_with_ema(store_params_in_ema=True)
    with torch.no_grad():
        assert O.rel_l2(netG(x, t, z).cpu(), y1.cpu()) < 1e-6


def test_flat_adam_checkpoint_surface_matches_torch_adam():
    """state_dict() / load_state_dict() in torch.optim.Adam's layout, per-network betas / weight decay, the LR-scheduler hook,
    and the EMA state in the reference's {name: cpu tensor} layout (ddgan.py:298-313, 545-569; ema.py:81-95)."""
    from ddgan_b200.train import FlatAdam
    torch.manual_seed(5)
    net_a = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 11)).to(DEV)
    net_b = copy.deepcopy(net_a)
    net_c = copy.deepcopy(net_a)
    fa = FlatAdam(net_a, 2e-4, (0.5, 0.999), weight_decay=1e-3, max_norm=1.0, ema_decay=0.99)
    opt = torch.optim.Adam(net_b.parameters(), lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    sched_a = torch.optim.lr_scheduler.CosineAnnealingLR(opt, 10, eta_min=1e-5)
    xs = [torch.randn(16, 37, device=DEV) * (5.0 if i % 2 else 0.05) for i in range(6)]
    for i, x in enumerate(xs[:3]):
        fa.zero_grad(); opt.zero_grad()
        (net_a(x) ** 2).sum().backward(); (net_b(x) ** 2).sum().backward()
        fa.set_lr(opt.param_groups[0]['lr'])
        fa.step()
        torch.nn.utils.clip_grad_norm_(net_b.parameters(), 1.0); opt.step(); sched_a.step()
    # checkpoint -> fresh FlatAdam on a copy of the current weights; torch's optimiser must accept the same dict
    sd = fa.state_dict()
    net_c.load_state_dict({k: v.detach().clone() for k, v in net_a.state_dict().items()})
    fc = FlatAdam(net_c, 1.0, (0.9, 0.9), max_norm=1.0, ema_decay=0.99)
    fc.load_state_dict(sd)
    fc.load_ema_state_dict(fa.ema_state_dict())
    torch.optim.Adam(copy.deepcopy(net_b).parameters(), lr=1.0).load_state_dict(sd)
    assert all(v.device.type == 'cpu' for v in fa.ema_state_dict().values())
    assert list(fa.ema_state_dict()) == [n for n, _ in net_a.named_parameters()]
    for i, x in enumerate(xs[3:]):
        lr = opt.param_groups[0]['lr']
        for f, net in ((fa, net_a), (fc, net_c)):
            f.zero_grad(); (net(x) ** 2).sum().backward(); f.set_lr(lr); f.step()
        opt.zero_grad(); (net_b(x) ** 2).sum().backward()
        torch.nn.utils.clip_grad_norm_(net_b.parameters(), 1.0); opt.step(); sched_a.step()
    for pa, pb, pc in zip(net_a.parameters(), net_b.parameters(), net_c.parameters()):
        assert O.rel_l2(pa.detach().cpu(), pb.detach().cpu()) < 1e-6
        assert O.rel_l2(pc.detach().cpu(), pa.detach().cpu()) < 1e-7
    assert O.rel_l2(fc.ema.cpu(), fa.ema.cpu()) < 1e-7


def test_trainer_reads_reference_optimizer_arguments():
    """ddgan.py:298-310 / train_ddgan.py:73-89: beta1_d/beta2_d/beta1_g/beta2_g and weight_decay_D/G, with the upstream
    beta1/beta2 = None as in the reference's argparse namespace."""
    from ddgan_b200.train import Trainer
    cfg, netG, netD = _tiny_nets()
    for k, v in dict(lr_g=1.6e-4, lr_d=1.25e-4, beta1=None, beta2=None, beta1_g=0.4, beta2_g=0.99, beta1_d=0.6, beta2_d=0.95,
                     weight_decay_G=1e-4, weight_decay_D=2e-4, r1_gamma=0.02, lazy_reg=15, grad_clip_norm=1.0, ema_decay=0.999,
                     use_ema=True).items():
        setattr(cfg, k, v)
    tr = Trainer(cfg, netG, netD, DEV)
    assert tr.optG.betas == (0.4, 0.99) and tr.optD.betas == (0.6, 0.95)
    assert tr.optG.wd == 1e-4 and tr.optD.wd == 2e-4
    real = torch.tanh(seeded((4, 3, 32, 32), 300)).to(DEV)
    eD, eG = tr.step(real, 0)
    assert torch.isfinite(eD) and torch.isfinite(eG)


def test_discriminator_large_r1_step_gradients_vs_oracle():
    """BASELINE configs[3] path: Discriminator_large on 256-px inputs, D-real loss + R1 penalty (gamma 1.0) with the double
    backward through six FIR-downsampling blocks, and the fake-sample loss; losses and every gradient tensor against CPU
    autograd through the oracle (ngf 8 keeps the CPU side fast; 16/32/64-channel maps = narrow-map and N = 64 wgrad paths).
    Tolerances as in the CIFAR step test: losses 1e-4, R1 5e-4, whole gradient 5e-4, live tensors 2e-3."""
    from ddgan_b200.modules import Discriminator_large
    B, ngf, ted, gamma = 2, 8, 32, 1.0
    shapes = O.discriminator_param_shapes(6, ngf, ted, large=True)
    sd = O.randomize_params(shapes, seed=71)
    x_t = seeded((B, 3, 256, 256), 710); x_tp1 = seeded((B, 3, 256, 256), 711); x_fake = seeded((B, 3, 256, 256), 712)
    t = torch.tensor([1, 3])
    # ---- oracle ----
    pd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x_t.clone().requires_grad_(True)
    d_real = O.discriminator_forward(pd, xr, t, x_tp1, ted, large=True).view(-1)
    er = F.softplus(-d_real).mean()
    g, = torch.autograd.grad(d_real.sum(), xr, create_graph=True)
    gp = gamma / 2 * (g.view(B, -1).norm(2, dim=1) ** 2).mean()
    ef = F.softplus(O.discriminator_forward(pd, x_fake, t, x_tp1, ted, large=True).view(-1)).mean()
    (er + gp + ef).backward()
    ref = {k: v.grad.clone() for k, v in pd.items()}
    # ---- this framework: ddgan.py:455-477 on the drop-in module ----
    net = Discriminator_large(nc=6, ngf=ngf, t_emb_dim=ted).to(DEV)
    net.load_state_dict(sd, strict=True)
    xd = x_t.to(DEV).requires_grad_(True); td = t.to(DEV)
    D_real = net(xd, td, x_tp1.to(DEV)).view(-1)
    errD_real = F.softplus(-D_real).mean()
    errD_real.backward(retain_graph=True)
    grad_real = torch.autograd.grad(outputs=D_real.sum(), inputs=xd, create_graph=True)[0]
    grad_penalty = gamma / 2 * (grad_real.view(B, -1).norm(2, dim=1) ** 2).mean()
    grad_penalty.backward()
    errD_fake = F.softplus(net(x_fake.to(DEV), td, x_tp1.to(DEV)).view(-1)).mean()
    errD_fake.backward()
    assert abs(float(errD_real) - float(er)) < 1e-4 * max(1.0, abs(float(er)))
    assert abs(float(errD_fake) - float(ef)) < 1e-4 * max(1.0, abs(float(ef)))
    assert abs(float(grad_penalty) - float(gp)) < 5e-4 * abs(float(gp))
    got = {k: p.grad.detach().cpu() for k, p in net.named_parameters()}
    keys = list(ref)
    flat = lambda d: torch.cat([d[k].flatten().double() for k in keys])
    total = float(flat(ref).norm())
    assert O.rel_l2(flat(got), flat(ref)) < 5e-4
    worst = max((O.rel_l2(got[k], ref[k]), k) for k in keys if float(ref[k].double().norm()) > 1e-6 * total)
    assert worst[0] < 2e-3, worst


def test_direct_gradient_accumulation_matches_autograd():
    """Trainer.step lets the wgrad / channel-sum kernels accumulate into the flat gradient arena (train_graph.ACCUM) and packs
    all conv weights with one launch per forward after the first step (TrainPacks); both must give what plain autograd
    accumulation with per-conv packing gives: same losses, both gradient arenas to 1e-5.  lr_d = 0 keeps the G step
    deterministic (see test_graphed_train_step_equals_eager_step)."""
    from ddgan_b200.train import Trainer
    from ddgan_b200 import train_graph as TG
    cfg, netG, netD = _tiny_nets()
    cfg = _train_cfg(cfg, lazy_reg=2)
    cfg.lr_d = 0.0
    netG2, netD2 = copy.deepcopy(netG), copy.deepcopy(netD)
    B = 4
    real = torch.tanh(seeded((B, 3, 32, 32), 300)).to(DEV)
    a = Trainer(cfg, netG, netD, DEV)
    b = Trainer(cfg, netG2, netD2, DEV)
    b._packs_frozen = True                      # never freeze: every conv packs its own operand, every call
    orig = TG._grad_target
    for it in range(3):                          # step 0 records the packs, steps 1-2 replay the plan (R1 on 0 and 2)
        nz = _noise(cfg, B, 32, 700 + 10 * it)
        ea = a.step(real, it, noise=nz)
        TG._grad_target = lambda p: None         # plain autograd: temporaries + AccumulateGrad
        try:
            eb = b.step(real, it, noise=nz)
        finally:
            TG._grad_target = orig
        assert abs(float(ea[0]) - float(eb[0])) < 1e-5 * max(1.0, abs(float(ea[0])))
        assert abs(float(ea[1]) - float(eb[1])) < 1e-5 * max(1.0, abs(float(ea[1])))
        assert O.rel_l2(a.optD.flat_g.cpu(), b.optD.flat_g.cpu()) < 1e-5, it
        assert O.rel_l2(a.optG.flat_g.cpu(), b.optG.flat_g.cpu()) < 1e-5, it
        _sync_state(a, b)
    assert TG.train_packs(netG).frozen and len(TG.train_packs(netG).plan.items) > 20
