"""CPU oracle for the DDGAN hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain functions over a flat
``state_dict``; no nn.Module, no CUDA) of the algorithms on the reference's hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it -- and there only as
the checker / the CPU baseline, never as the product path.  The product path
(``denoising-diffusion-gan_b200/``) never imports anything from ``oracle/``.

Parity pin: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  The oracle is pinned instead against the reference
itself, imported unmodified from ``/root/reference`` in the build container:
``tests/golden/make_golden.py`` runs the reference's own ``upfirdn2d_native``,
``NCSNpp``, ``Discriminator_small/large``, schedules and sampler on seeded
inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against them.

Third-party arithmetic: conv / linear / softmax / erf-free activations are
PyTorch CPU kernels (torch 2.11.0, the version pinned by this image), exactly
as in the reference's CPU path.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

RSQRT2 = 1.0 / math.sqrt(2.0)

# ----------------------------------------------------------------------------
# L0/L1: upfirdn2d and fused bias+leaky-relu
# ----------------------------------------------------------------------------


def setup_fir_kernel(k) -> np.ndarray:
    """score_sde/models/up_or_down_sampling.py:186-193 (_setup_kernel):
    separable taps -> outer product, normalised to sum 1, float32."""
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k = k / np.sum(k)
    assert k.ndim == 2 and k.shape[0] == k.shape[1]
    return k.astype(np.float32)


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up: int = 1, down: int = 1,
              pad=(0, 0)) -> torch.Tensor:
    """score_sde/op/upfirdn2d.py:153-164 + 184-225 (upfirdn2d_native semantics).

    Direct statement of the definition, tap by tap (no conv2d call):
      1. zero-insert:   u[y*up, x*up] = in[y, x]
      2. pad (negative pad crops) by (pad0, pad1) on both axes
      3. out_full[y, x] = sum_{i,j} u_pad[y+i, x+j] * kernel[kh-1-i, kw-1-j]
         (true convolution = correlation with the flipped kernel, :211-212)
      4. keep every `down`-th sample.
    """
    n, c, in_h, in_w = x.shape
    kh, kw = kernel.shape
    p0, p1 = pad
    u = x.new_zeros(n, c, in_h * up, in_w * up)
    u[:, :, ::up, ::up] = x
    u = F.pad(u, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    u = u[:, :, max(-p0, 0): u.shape[2] - max(-p1, 0), max(-p0, 0): u.shape[3] - max(-p1, 0)]
    full_h = in_h * up + p0 + p1 - kh + 1
    full_w = in_w * up + p0 + p1 - kw + 1
    out = x.new_zeros(n, c, full_h, full_w)
    for i in range(kh):
        for j in range(kw):
            out = out + u[:, :, i:i + full_h, j:j + full_w] * kernel[kh - 1 - i, kw - 1 - j]
    return out[:, :, ::down, ::down].contiguous()


def fused_leaky_relu(x: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2,
                     scale: float = 2 ** 0.5) -> torch.Tensor:
    """score_sde/op/fused_act.py:94-105 and fused_bias_act_kernel.cu:20-51 (act*10+grad == 30).

    NOTE: the reference's *CPU* branch hard-codes slope 0.2 (fused_act.py:99) while its CUDA
    kernel honours `negative_slope`; the oracle follows the CUDA kernel (the path we replace)."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    y = x + bias.view(*shape)
    return torch.where(y > 0, y, y * negative_slope) * scale


def fused_leaky_relu_grad(grad_out: torch.Tensor, out: torch.Tensor, negative_slope: float = 0.2,
                          scale: float = 2 ** 0.5):
    """score_sde/op/fused_act.py:28-50 (FusedLeakyReLUFunctionBackward.forward; kernel case 31):
    grad_in = (out > 0 ? g : g*alpha) * scale ; grad_bias = sum over all dims but 1."""
    gi = torch.where(out > 0, grad_out, grad_out * negative_slope) * scale
    dims = [0] + list(range(2, gi.ndim))
    return gi, gi.sum(dims)


def upsample_2d(x, k=(1, 3, 3, 1), factor: int = 2, gain: float = 1.0):
    """score_sde/models/up_or_down_sampling.py:200-228."""
    kk = setup_fir_kernel(k) * (gain * factor ** 2)
    p = kk.shape[0] - factor
    return upfirdn2d(x, torch.from_numpy(kk).to(x), up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k=(1, 3, 3, 1), factor: int = 2, gain: float = 1.0):
    """score_sde/models/up_or_down_sampling.py:231-261."""
    kk = setup_fir_kernel(k) * gain
    p = kk.shape[0] - factor
    return upfirdn2d(x, torch.from_numpy(kk).to(x), down=factor, pad=((p + 1) // 2, p // 2))


def conv_downsample_2d(x, w, k=(1, 3, 3, 1), factor: int = 2, gain: float = 1.0):
    """score_sde/models/up_or_down_sampling.py:149-183: FIR with pad (2,2) then stride-2 VALID conv."""
    kk = setup_fir_kernel(k) * gain
    conv_w = w.shape[-1]
    p = (kk.shape[0] - factor) + (conv_w - 1)
    x = upfirdn2d(x, torch.from_numpy(kk).to(x), pad=((p + 1) // 2, p // 2))
    return F.conv2d(x, w, stride=factor, padding=0)


# ----------------------------------------------------------------------------
# L2 building blocks
# ----------------------------------------------------------------------------


def timestep_embedding(t: torch.Tensor, dim: int, max_positions: int = 10000) -> torch.Tensor:
    """score_sde/models/layers.py:475-486: [sin | cos] of t * exp(-j ln(max)/(half-1))."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -(math.log(max_positions) / (half - 1)))
    arg = t.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def group_norm(x: torch.Tensor, groups: int, eps: float = 1e-6, weight=None, bias=None) -> torch.Tensor:
    """nn.GroupNorm semantics (biased variance over (C/G, H, W)), written out explicitly.
    Used at layerspp.py:50,100 and ncsnpp_generator_adagn.py:264."""
    n, c, h, w = x.shape
    xg = x.reshape(n, groups, -1).double()
    mean = xg.mean(dim=2, keepdim=True)
    var = ((xg - mean) ** 2).mean(dim=2, keepdim=True)
    y = ((xg - mean) / torch.sqrt(var + eps)).to(x.dtype).reshape(n, c, h, w)
    if weight is not None:
        y = y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    return y


def num_groups(c: int) -> int:
    """layerspp.py:254,267 / :100: G = min(C // 4, 32)."""
    return min(c // 4, 32)


def adaptive_group_norm(x, zemb, style_w, style_b):
    """score_sde/models/layerspp.py:46-63: [gamma|beta] = Linear(zemb); gamma*GN(x)+beta."""
    c = x.shape[1]
    style = F.linear(zemb, style_w, style_b)
    gamma, beta = style[:, :c, None, None], style[:, c:, None, None]
    return gamma * group_norm(x, num_groups(c)) + beta


def nin(x, W, b):
    """score_sde/models/layers.py:489-512: 1x1 'network in network' with weight stored (in, out)."""
    return torch.einsum('bchw,co->bohw', x, W) + b.view(1, -1, 1, 1)


def attn_block(x, sd, prefix, skip_rescale=True):
    """score_sde/models/layerspp.py:95-124 (AttnBlockpp.forward)."""
    n, c, h, w = x.shape
    g = group_norm(x, num_groups(c), weight=sd[prefix + 'GroupNorm_0.weight'], bias=sd[prefix + 'GroupNorm_0.bias'])
    q = nin(g, sd[prefix + 'NIN_0.W'], sd[prefix + 'NIN_0.b']).reshape(n, c, h * w)
    k = nin(g, sd[prefix + 'NIN_1.W'], sd[prefix + 'NIN_1.b']).reshape(n, c, h * w)
    v = nin(g, sd[prefix + 'NIN_2.W'], sd[prefix + 'NIN_2.b']).reshape(n, c, h * w)
    logits = torch.einsum('bcq,bck->bqk', q, k) * (int(c) ** (-0.5))
    p = torch.softmax(logits, dim=-1)
    o = torch.einsum('bqk,bck->bcq', p, v).reshape(n, c, h, w)
    o = nin(o, sd[prefix + 'NIN_3.W'], sd[prefix + 'NIN_3.b'])
    return (x + o) * RSQRT2 if skip_rescale else x + o


def resblock_biggan(x, temb, zemb, sd, prefix, up=False, down=False, fir_kernel=(1, 3, 3, 1), skip_rescale=True):
    """score_sde/models/layerspp.py:247-310 (ResnetBlockBigGANpp_Adagn.forward), dropout = identity (eval
    or p = 0, the README configs)."""
    h = F.silu(adaptive_group_norm(x, zemb, sd[prefix + 'GroupNorm_0.style.weight'], sd[prefix + 'GroupNorm_0.style.bias']))
    if up:
        h = upsample_2d(h, fir_kernel)
        x = upsample_2d(x, fir_kernel)
    elif down:
        h = downsample_2d(h, fir_kernel)
        x = downsample_2d(x, fir_kernel)
    h = F.conv2d(h, sd[prefix + 'Conv_0.weight'], sd[prefix + 'Conv_0.bias'], padding=1)
    if temb is not None:
        h = h + F.linear(F.silu(temb), sd[prefix + 'Dense_0.weight'], sd[prefix + 'Dense_0.bias'])[:, :, None, None]
    h = F.silu(adaptive_group_norm(h, zemb, sd[prefix + 'GroupNorm_1.style.weight'], sd[prefix + 'GroupNorm_1.style.bias']))
    h = F.conv2d(h, sd[prefix + 'Conv_1.weight'], sd[prefix + 'Conv_1.bias'], padding=1)
    if (prefix + 'Conv_2.weight') in sd:
        x = F.conv2d(x, sd[prefix + 'Conv_2.weight'], sd[prefix + 'Conv_2.bias'])
    return (x + h) * RSQRT2 if skip_rescale else x + h


# ----------------------------------------------------------------------------
# NCSN++ generator (biggan blocks, positional embedding, progressive_input residual|none)
# ----------------------------------------------------------------------------


def cifar10_config(**over) -> SimpleNamespace:
    """BASELINE.json configs[0]/[1]; SURVEY.md section 8 header (readme.md:31-37)."""
    cfg = dict(image_size=32, num_channels=3, num_channels_dae=128, ch_mult=(1, 2, 2, 2), num_res_blocks=2,
               attn_resolutions=(16,), dropout=0.0, resamp_with_conv=True, conditional=True, fir=True,
               fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type='biggan', progressive='none',
               progressive_input='residual', progressive_combine='sum', embedding_type='positional',
               fourier_scale=16.0, not_use_tanh=False, z_emb_dim=256, nz=100, n_mlp=4, centered=True,
               t_emb_dim=256, ngf=64, num_timesteps=4, beta_min=0.1, beta_max=20.0, use_geometric=False)
    cfg.update(over)
    return SimpleNamespace(**cfg)


def tiny_config(**over) -> SimpleNamespace:
    """A shrunken config with the same topology (attention level, pyramid, up/down blocks) for fast CPU tests."""
    return cifar10_config(**{**dict(image_size=16, num_channels_dae=32, ch_mult=(1, 2, 2), attn_resolutions=(8,),
                                    z_emb_dim=64, nz=20, n_mlp=2, t_emb_dim=32, ngf=16), **over})


def celebahq256_config(**over) -> SimpleNamespace:
    """BASELINE.json configs[2]/[3] (readme.md:41-55)."""
    return cifar10_config(**{**dict(image_size=256, num_channels_dae=64, ch_mult=(1, 1, 2, 2, 4, 4), n_mlp=3,
                                    num_timesteps=2, ngf=64), **over})


def ncsnpp_forward(sd: dict, cfg, x: torch.Tensor, t: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """score_sde/models/ncsnpp_generator_adagn.py:280-431 for resblock_type='biggan',
    embedding_type='positional', progressive='none', progressive_input in {'residual','none'}.
    Module indices follow the construction order at :93-267."""
    assert cfg.resblock_type == 'biggan' and cfg.embedding_type == 'positional' and cfg.progressive == 'none'
    nf = cfg.num_channels_dae
    nres = len(cfg.ch_mult)
    fk = tuple(cfg.fir_kernel)
    # z mapping network, :51-56 and :271-277
    zn = z / torch.sqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    zemb = F.silu(F.linear(zn, sd['z_transform.1.weight'], sd['z_transform.1.bias']))
    for i in range(cfg.n_mlp):
        zemb = F.silu(F.linear(zemb, sd[f'z_transform.{3 + 2 * i}.weight'], sd[f'z_transform.{3 + 2 * i}.bias']))
    # time embedding, :295-304
    m = 0
    temb = timestep_embedding(t, nf)
    if cfg.conditional:
        temb = F.linear(temb, sd['all_modules.0.weight'], sd['all_modules.0.bias'])
        temb = F.linear(F.silu(temb), sd['all_modules.1.weight'], sd['all_modules.1.bias'])
        m = 2
    else:
        temb = None
    if not cfg.centered:
        x = 2 * x - 1.0
    pyramid = x if cfg.progressive_input != 'none' else None

    def P(i):
        return f'all_modules.{i}.'

    hs = [F.conv2d(x, sd[P(m) + 'weight'], sd[P(m) + 'bias'], padding=1)]
    m += 1
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            h = resblock_biggan(hs[-1], temb, zemb, sd, P(m), fir_kernel=fk, skip_rescale=cfg.skip_rescale)
            m += 1
            if h.shape[-1] in cfg.attn_resolutions:
                h = attn_block(h, sd, P(m), cfg.skip_rescale)
                m += 1
            hs.append(h)
        if lvl != nres - 1:
            h = resblock_biggan(hs[-1], temb, zemb, sd, P(m), down=True, fir_kernel=fk, skip_rescale=cfg.skip_rescale)
            m += 1
            if cfg.progressive_input == 'residual':
                # layerspp.py:194 -> up_or_down_sampling.py:52-59 (conv_downsample_2d + bias)
                pyramid = conv_downsample_2d(pyramid, sd[P(m) + 'Conv2d_0.weight'], fk) \
                    + sd[P(m) + 'Conv2d_0.bias'].view(1, -1, 1, 1)
                m += 1
                pyramid = (pyramid + h) * RSQRT2 if cfg.skip_rescale else pyramid + h
                h = pyramid
            hs.append(h)
    h = hs[-1]
    h = resblock_biggan(h, temb, zemb, sd, P(m), fir_kernel=fk, skip_rescale=cfg.skip_rescale); m += 1
    h = attn_block(h, sd, P(m), cfg.skip_rescale); m += 1
    h = resblock_biggan(h, temb, zemb, sd, P(m), fir_kernel=fk, skip_rescale=cfg.skip_rescale); m += 1
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            h = resblock_biggan(torch.cat([h, hs.pop()], dim=1), temb, zemb, sd, P(m), fir_kernel=fk,
                                skip_rescale=cfg.skip_rescale)
            m += 1
        if h.shape[-1] in cfg.attn_resolutions:
            h = attn_block(h, sd, P(m), cfg.skip_rescale)
            m += 1
        if lvl != 0:
            h = resblock_biggan(h, temb, zemb, sd, P(m), up=True, fir_kernel=fk, skip_rescale=cfg.skip_rescale)
            m += 1
    assert not hs
    c = h.shape[1]
    h = F.silu(group_norm(h, num_groups(c), weight=sd[P(m) + 'weight'], bias=sd[P(m) + 'bias'])); m += 1
    h = F.conv2d(h, sd[P(m) + 'weight'], sd[P(m) + 'bias'], padding=1); m += 1
    return h if cfg.not_use_tanh else torch.tanh(h)


# ----------------------------------------------------------------------------
# Discriminators
# ----------------------------------------------------------------------------


def _leaky(x):
    return F.leaky_relu(x, 0.2)


def down_conv_block(x, t_emb, sd, prefix, downsample, fir_kernel=(1, 3, 3, 1)):
    """score_sde/models/discriminator.py:76-94 (DownConvBlock.forward)."""
    out = F.conv2d(_leaky(x), sd[prefix + 'conv1.0.weight'], sd[prefix + 'conv1.0.bias'], padding=1)
    out = out + F.linear(t_emb, sd[prefix + 'dense_t1.weight'], sd[prefix + 'dense_t1.bias'])[..., None, None]
    out = _leaky(out)
    if downsample:
        out = downsample_2d(out, fir_kernel)
        x = downsample_2d(x, fir_kernel)
    out = F.conv2d(out, sd[prefix + 'conv2.0.weight'], sd[prefix + 'conv2.0.bias'], padding=1)
    skip = F.conv2d(x, sd[prefix + 'skip.0.weight'])
    return (out + skip) * RSQRT2


def discriminator_forward(sd: dict, x, t, x_t, t_emb_dim: int, large: bool = False, stages: list = None):
    """score_sde/models/discriminator.py:134-167 (small) / :205-238 (large).  `stages` (test aid): a list that receives the
    output of start_conv, of every DownConvBlock, of final_conv (after the activation) and the pooled feature vector.
    Runs in the dtype of `x` (float64 inputs and weights give the float64 ground truth; the sinusoid table stays the
    reference's float32 one)."""
    te = timestep_embedding(t, t_emb_dim).to(x.dtype)
    te = F.linear(te, sd['t_embed.main.0.weight'], sd['t_embed.main.0.bias'])
    te = F.linear(_leaky(te), sd['t_embed.main.2.weight'], sd['t_embed.main.2.bias'])
    te = _leaky(te)
    h = F.conv2d(torch.cat((x, x_t), dim=1), sd['start_conv.weight'], sd['start_conv.bias'])
    if stages is not None:
        stages.append(h)
    if large:
        flags = [True] * 6
    else:
        flags = [False, True, True, True]
    for i, ds in enumerate(flags):
        h = down_conv_block(h, te, sd, f'conv{i + 1}.', ds)
        if stages is not None:
            stages.append(h)
    b, c, hh, ww = h.shape
    group = min(b, 4)
    # minibatch stddev, :150-158: sample i is grouped with i + B/group, i + 2B/group, ...
    g = h.view(group, -1, 1, c, hh, ww)
    sdv = torch.sqrt(g.var(0, unbiased=False) + 1e-8).mean([2, 3, 4], keepdim=True).squeeze(2)
    sdv = sdv.repeat(group, 1, hh, ww)
    h = torch.cat([h, sdv], 1)
    h = _leaky(F.conv2d(h, sd['final_conv.weight'], sd['final_conv.bias'], padding=1))
    if stages is not None:
        stages.append(h)
    h = h.view(b, h.shape[1], -1).sum(2)
    if stages is not None:
        stages.append(h)
    return F.linear(h, sd['end_linear.weight'], sd['end_linear.bias'])


# ----------------------------------------------------------------------------
# L3 diffusion math
# ----------------------------------------------------------------------------


def sigma_schedule(num_timesteps: int, beta_min: float, beta_max: float, use_geometric: bool = False):
    """ddgan.py:38-90 / test_ddgan.py:11-63: tables of length T+1 computed in float64, betas cast to float32."""
    eps_small = 1e-3
    t = np.arange(0, num_timesteps + 1, dtype=np.float64) / num_timesteps
    t = torch.from_numpy(t) * (1.0 - eps_small) + eps_small
    if use_geometric:
        var = beta_min * ((beta_max / beta_min) ** t)
    else:
        var = 1.0 - torch.exp(2.0 * (-0.25 * t ** 2 * (beta_max - beta_min) - 0.5 * t * beta_min))
    alpha_bars = 1.0 - var
    betas = 1 - alpha_bars[1:] / alpha_bars[:-1]
    betas = torch.cat((torch.tensor([1e-8], dtype=torch.float64), betas)).type(torch.float32)
    return betas ** 0.5, torch.sqrt(1 - betas), betas


def diffusion_coefficients(cfg):
    """ddgan.py:93-107 (DiffusionCoefficients)."""
    sigmas, a_s, _ = sigma_schedule(cfg.num_timesteps, cfg.beta_min, cfg.beta_max, cfg.use_geometric)
    a_s_cum = torch.from_numpy(np.cumprod(a_s.numpy()))
    sigmas_cum = torch.sqrt(1 - a_s_cum ** 2)
    a_s_prev = a_s.clone()
    a_s_prev[-1] = 1
    return SimpleNamespace(sigmas=sigmas, a_s=a_s, a_s_cum=a_s_cum, sigmas_cum=sigmas_cum, a_s_prev=a_s_prev)


def posterior_coefficients(cfg):
    """ddgan.py:131-149 / test_ddgan.py:67-93."""
    _, _, betas = sigma_schedule(cfg.num_timesteps, cfg.beta_min, cfg.beta_max, cfg.use_geometric)
    betas = betas.type(torch.float32)[1:]
    alphas = 1 - betas
    ac = torch.cumprod(alphas, 0)
    ac_prev = torch.cat((torch.tensor([1.0]), ac[:-1]), 0)
    post_var = betas * (1 - ac_prev) / (1 - ac)
    return SimpleNamespace(
        betas=betas, alphas=alphas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev, posterior_variance=post_var,
        posterior_mean_coef1=betas * torch.sqrt(ac_prev) / (1 - ac),
        posterior_mean_coef2=(1 - ac_prev) * torch.sqrt(alphas) / (1 - ac),
        posterior_log_variance_clipped=torch.log(post_var.clamp(min=1e-20)))


def _per_sample(table, t, ndim):
    """ddgan.py:51-56 (extract)."""
    return table[t].reshape(-1, *([1] * (ndim - 1)))


def q_sample_pairs(coeff, x0, t, noise_xt, noise_xtp1):
    """ddgan.py:110-126.  The reference draws `noise` for x_{t+1} first (:122) and the q_sample noise second
    (:112); here both are injected: noise_xt feeds x_t, noise_xtp1 feeds x_{t+1}."""
    x_t = _per_sample(coeff.a_s_cum, t, x0.ndim) * x0 + _per_sample(coeff.sigmas_cum, t, x0.ndim) * noise_xt
    x_tp1 = _per_sample(coeff.a_s, t + 1, x0.ndim) * x_t + _per_sample(coeff.sigmas, t + 1, x0.ndim) * noise_xtp1
    return x_t, x_tp1


def sample_posterior(pc, x0, x_t, t, noise):
    """ddgan.py:152-169 / test_ddgan.py:96-113 with the randn_like draw injected."""
    mean = _per_sample(pc.posterior_mean_coef1, t, x_t.ndim) * x0 + _per_sample(pc.posterior_mean_coef2, t, x_t.ndim) * x_t
    log_var = _per_sample(pc.posterior_log_variance_clipped, t, x_t.ndim)
    mask = (t != 0).float().reshape(-1, *([1] * (x_t.ndim - 1)))
    return mean + mask * torch.exp(0.5 * log_var) * noise


def sample_from_model(pc, generator, n_time, x_init, nz, noise_fn=None):
    """ddgan.py:172-183 / test_ddgan.py:116-125.  `generator(x, t, z)`; `noise_fn(shape)` supplies the two
    normal draws per step in the reference's order (latent z, then posterior noise); default torch.randn."""
    noise_fn = noise_fn or (lambda shape: torch.randn(*shape))
    x = x_init
    with torch.no_grad():
        for i in reversed(range(n_time)):
            t = torch.full((x.size(0),), i, dtype=torch.int64)
            z = noise_fn((x.size(0), nz))
            x0 = generator(x, t, z)
            x = sample_posterior(pc, x0, x, t, noise_fn(tuple(x.shape))).detach()
    return x


# ----------------------------------------------------------------------------
# Train step body (losses + gradients), ddgan.py:443-518, with injected randomness
# ----------------------------------------------------------------------------


def d_step_losses(sd_g, sd_d, cfg, real, t, noises, z, r1_gamma, do_r1, large=False):
    """ddgan.py:449-477.  Returns (errD_real, grad_penalty or None, errD_fake) as graph-attached scalars.
    `noises` = (noise_xt, noise_xtp1, posterior_noise)."""
    coeff, pc = diffusion_coefficients(cfg), posterior_coefficients(cfg)
    x_t, x_tp1 = q_sample_pairs(coeff, real, t, noises[0], noises[1])
    x_t = x_t.detach().requires_grad_(True)
    d_real = discriminator_forward(sd_d, x_t, t, x_tp1.detach(), cfg.t_emb_dim, large).view(-1)
    err_real = F.softplus(-d_real).mean()
    gp = None
    if do_r1:
        g, = torch.autograd.grad(d_real.sum(), x_t, create_graph=True)
        gp = r1_gamma / 2 * (g.view(g.size(0), -1).norm(2, dim=1) ** 2).mean()
    x0p = ncsnpp_forward(sd_g, cfg, x_tp1.detach(), t, z)
    x_pos = sample_posterior(pc, x0p, x_tp1, t, noises[2])
    out = discriminator_forward(sd_d, x_pos, t, x_tp1.detach(), cfg.t_emb_dim, large).view(-1)
    err_fake = F.softplus(out).mean()
    return err_real, gp, err_fake


def g_step_loss(sd_g, sd_d, cfg, real, t, noises, z, large=False):
    """ddgan.py:495-503."""
    coeff, pc = diffusion_coefficients(cfg), posterior_coefficients(cfg)
    _, x_tp1 = q_sample_pairs(coeff, real, t, noises[0], noises[1])
    x0p = ncsnpp_forward(sd_g, cfg, x_tp1.detach(), t, z)
    x_pos = sample_posterior(pc, x0p, x_tp1, t, noises[2])
    out = discriminator_forward(sd_d, x_pos, t, x_tp1.detach(), cfg.t_emb_dim, large).view(-1)
    return F.softplus(-out).mean()


# ----------------------------------------------------------------------------
# Parameter factories (shapes/names = the reference's state_dict; values seeded N(0, sigma))
# ----------------------------------------------------------------------------


def ncsnpp_param_shapes(cfg) -> dict:
    """Names and shapes of NCSNpp(cfg).state_dict() in construction order
    (ncsnpp_generator_adagn.py:93-277, layerspp.py:247-276, :95-106, :166-185)."""
    nf, zd = cfg.num_channels_dae, cfg.z_emb_dim
    shapes = {}
    m = 0

    def lin(name, i, o):
        shapes[name + '.weight'] = (o, i); shapes[name + '.bias'] = (o,)

    def conv(name, i, o, k):
        shapes[name + '.weight'] = (o, i, k, k); shapes[name + '.bias'] = (o,)

    def block(i, o, resample=False):
        nonlocal m
        p = f'all_modules.{m}.'
        lin(p + 'GroupNorm_0.style', zd, 2 * i)
        conv(p + 'Conv_0', i, o, 3)
        lin(p + 'Dense_0', 4 * nf, o)
        lin(p + 'GroupNorm_1.style', zd, 2 * o)
        conv(p + 'Conv_1', o, o, 3)
        if i != o or resample:
            conv(p + 'Conv_2', i, o, 1)
        m += 1

    def attn(c):
        nonlocal m
        p = f'all_modules.{m}.'
        shapes[p + 'GroupNorm_0.weight'] = (c,); shapes[p + 'GroupNorm_0.bias'] = (c,)
        for j in range(4):
            shapes[p + f'NIN_{j}.W'] = (c, c); shapes[p + f'NIN_{j}.b'] = (c,)
        m += 1

    if cfg.conditional:
        lin('all_modules.0', nf, 4 * nf); lin('all_modules.1', 4 * nf, 4 * nf); m = 2
    conv(f'all_modules.{m}', cfg.num_channels, nf, 3); m += 1
    nres = len(cfg.ch_mult)
    res = [cfg.image_size // 2 ** i for i in range(nres)]
    hs_c = [nf]
    in_ch = nf
    pyr_ch = cfg.num_channels
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            out_ch = nf * cfg.ch_mult[lvl]
            block(in_ch, out_ch); in_ch = out_ch
            if res[lvl] in cfg.attn_resolutions:
                attn(in_ch)
            hs_c.append(in_ch)
        if lvl != nres - 1:
            block(in_ch, in_ch, resample=True)
            if cfg.progressive_input == 'residual':
                conv(f'all_modules.{m}.Conv2d_0', pyr_ch, in_ch, 3); m += 1
                pyr_ch = in_ch
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    block(in_ch, in_ch); attn(in_ch); block(in_ch, in_ch)
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            out_ch = nf * cfg.ch_mult[lvl]
            block(in_ch + hs_c.pop(), out_ch); in_ch = out_ch
        if res[lvl] in cfg.attn_resolutions:
            attn(in_ch)
        if lvl != 0:
            block(in_ch, in_ch, resample=True)
    shapes[f'all_modules.{m}.weight'] = (in_ch,); shapes[f'all_modules.{m}.bias'] = (in_ch,); m += 1
    conv(f'all_modules.{m}', in_ch, cfg.num_channels, 3); m += 1
    lin('z_transform.1', cfg.nz, zd)
    for i in range(cfg.n_mlp):
        lin(f'z_transform.{3 + 2 * i}', zd, zd)
    return shapes


def discriminator_param_shapes(nc: int, ngf: int, t_emb_dim: int, large: bool = False) -> dict:
    """Names/shapes of Discriminator_small|large.state_dict() (discriminator.py:96-132, :170-203)."""
    shapes = {}

    def lin(name, i, o):
        shapes[name + '.weight'] = (o, i); shapes[name + '.bias'] = (o,)

    lin('t_embed.main.0', t_emb_dim, t_emb_dim); lin('t_embed.main.2', t_emb_dim, t_emb_dim)
    shapes['start_conv.weight'] = (ngf * 2, nc, 1, 1); shapes['start_conv.bias'] = (ngf * 2,)
    if large:
        chans = [(2, 4), (4, 8), (8, 8), (8, 8), (8, 8), (8, 8)]
    else:
        chans = [(2, 2), (2, 4), (4, 8), (8, 8)]
    for i, (a, b) in enumerate(chans):
        p = f'conv{i + 1}.'
        shapes[p + 'conv1.0.weight'] = (ngf * b, ngf * a, 3, 3); shapes[p + 'conv1.0.bias'] = (ngf * b,)
        shapes[p + 'conv2.0.weight'] = (ngf * b, ngf * b, 3, 3); shapes[p + 'conv2.0.bias'] = (ngf * b,)
        lin(p + 'dense_t1', t_emb_dim, ngf * b)
        shapes[p + 'skip.0.weight'] = (ngf * b, ngf * a, 1, 1)
    shapes['final_conv.weight'] = (ngf * 8, ngf * 8 + 1, 3, 3); shapes['final_conv.bias'] = (ngf * 8,)
    lin('end_linear', ngf * 8, 1)
    return shapes


def randomize_params(shapes: dict, seed: int, gain: float = 1.0) -> dict:
    """Re-randomised weights for meaningful parity (SURVEY.md 'five things' #5: reference init scales Conv_1 /
    NIN_3 / output conv by 1e-10, hiding errors).  Weights ~ N(0, gain/fan_in), biases ~ N(0, 0.1^2); AdaGN style
    biases get the reference's (1, 0) offsets (layerspp.py:53-54); GroupNorm affine weights ~ 1 + 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shp in shapes.items():
        if name.endswith('.W'):  # NIN (in, out)
            fan_in = shp[0]
            sd[name] = torch.randn(shp, generator=g) * math.sqrt(gain / fan_in)
        elif len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            sd[name] = torch.randn(shp, generator=g) * math.sqrt(gain / fan_in)
        else:
            v = torch.randn(shp, generator=g) * 0.1
            if name.endswith('style.bias'):
                v[: shp[0] // 2] += 1.0
            elif 'GroupNorm_0.weight' in name or (name.endswith('.weight') and len(shp) == 1):
                v += 1.0
            sd[name] = v
    return sd


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """Relative L2 error ||a-b|| / ||b|| in float64 (the north-star parity metric)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
