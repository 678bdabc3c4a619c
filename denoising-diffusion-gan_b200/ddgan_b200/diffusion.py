"""Diffusion schedules and the Gaussian updates of the DDGAN process (ddgan.py:38-183, test_ddgan.py:11-125).

The coefficient tables have T+1 (or T) entries and are computed once on the host in float64 exactly as the reference
does (then cast to float32); the per-sample updates run as single fused kernels (ddg_q_sample_pairs,
ddg_sample_posterior)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def get_time_schedule(num_timesteps, device='cpu'):
    """ddgan.py:59-66."""
    eps_small = 1e-3
    t = np.arange(0, num_timesteps + 1, dtype=np.float64) / num_timesteps
    return (torch.from_numpy(t) * (1. - eps_small) + eps_small).to(device)


def get_sigma_schedule(num_timesteps, beta_min, beta_max, use_geometric=False):
    """ddgan.py:69-90: returns (sigmas, a_s, betas), float32, length T+1."""
    t = get_time_schedule(num_timesteps)
    if use_geometric:
        var = beta_min * ((beta_max / beta_min) ** t)
    else:
        var = 1. - torch.exp(2. * (-0.25 * t ** 2 * (beta_max - beta_min) - 0.5 * t * beta_min))
    alpha_bars = 1.0 - var
    betas = 1 - alpha_bars[1:] / alpha_bars[:-1]
    betas = torch.cat((torch.tensor([1e-8], dtype=torch.float64), betas)).type(torch.float32)
    return betas ** 0.5, torch.sqrt(1 - betas), betas


class DiffusionCoefficients:
    """ddgan.py:93-107."""

    def __init__(self, args, device):
        sigmas, a_s, _ = get_sigma_schedule(args.num_timesteps, args.beta_min, args.beta_max, getattr(args, 'use_geometric', False))
        a_s_cum = torch.from_numpy(np.cumprod(a_s.numpy()))
        sigmas_cum = torch.sqrt(1 - a_s_cum ** 2)
        a_s_prev = a_s.clone()
        a_s_prev[-1] = 1
        self.sigmas, self.a_s = sigmas.to(device), a_s.to(device)
        self.a_s_cum, self.sigmas_cum, self.a_s_prev = a_s_cum.to(device), sigmas_cum.to(device), a_s_prev.to(device)


class PosteriorCoefficients:
    """ddgan.py:131-149 / test_ddgan.py:67-93 (Posterior_Coefficients)."""

    def __init__(self, args, device):
        _, _, betas = get_sigma_schedule(args.num_timesteps, args.beta_min, args.beta_max, getattr(args, 'use_geometric', False))
        betas = betas.type(torch.float32)[1:]
        alphas = 1 - betas
        ac = torch.cumprod(alphas, 0)
        ac_prev = torch.cat((torch.tensor([1.], dtype=torch.float32), ac[:-1]), 0)
        pv = betas * (1 - ac_prev) / (1 - ac)
        self.betas, self.alphas = betas.to(device), alphas.to(device)
        self.alphas_cumprod, self.alphas_cumprod_prev = ac.to(device), ac_prev.to(device)
        self.posterior_variance = pv.to(device)
        self.sqrt_alphas_cumprod = torch.sqrt(ac).to(device)
        self.sqrt_recip_alphas_cumprod = torch.rsqrt(ac).to(device)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1 / ac - 1).to(device)
        self.posterior_mean_coef1 = (betas * torch.sqrt(ac_prev) / (1 - ac)).to(device)
        self.posterior_mean_coef2 = ((1 - ac_prev) * torch.sqrt(alphas) / (1 - ac)).to(device)
        self.posterior_log_variance_clipped = torch.log(pv.clamp(min=1e-20)).to(device)


Posterior_Coefficients = PosteriorCoefficients  # test_ddgan.py spelling


def q_sample_pairs(coeff, x_start, t, noise_xt=None, noise_xtp1=None):
    """ddgan.py:119-126.  Draw order of the reference: x_{t+1} noise first (:122), then the q_sample noise (:112)."""
    if noise_xtp1 is None:
        noise_xtp1 = torch.randn_like(x_start)
    if noise_xt is None:
        noise_xt = torch.randn_like(x_start)
    return ops.q_sample_pairs(x_start, noise_xt, noise_xtp1, t, coeff.a_s_cum, coeff.sigmas_cum, coeff.a_s, coeff.sigmas)


def sample_posterior(coefficients, x_0, x_t, t, noise=None, out=None):
    """ddgan.py:152-169.  Under autograd (training: x_0 = G(...) carries a graph) the three-term update is expressed with
    differentiable elementwise ops; otherwise it is the single fused kernel."""
    if noise is None:
        noise = torch.randn_like(x_t)
    if torch.is_grad_enabled() and (x_0.requires_grad or x_t.requires_grad):
        shp = (-1,) + (1,) * (x_t.dim() - 1)
        mean = coefficients.posterior_mean_coef1[t].view(shp) * x_0 + coefficients.posterior_mean_coef2[t].view(shp) * x_t
        sd = (t != 0).float().view(shp) * torch.exp(0.5 * coefficients.posterior_log_variance_clipped[t].view(shp))
        return mean + sd * noise
    return ops.sample_posterior(x_0, x_t, noise, t, coefficients.posterior_mean_coef1, coefficients.posterior_mean_coef2,
                                coefficients.posterior_log_variance_clipped, out=out)


def sample_from_model(coefficients, generator, n_time, x_init, T, opt, noise_fn=None):
    """ddgan.py:172-183 / test_ddgan.py:116-125; `generator(x, t, z)`.  noise_fn(shape) overrides torch.randn for parity
    runs (the two draws per step come in the reference's order: latent z, then the posterior noise)."""
    dev = x_init.device
    noise_fn = noise_fn or (lambda shape: torch.randn(*shape, device=dev))
    x = x_init
    with torch.no_grad():
        for i in reversed(range(n_time)):
            t = torch.full((x.size(0),), i, dtype=torch.int64, device=dev)
            z = noise_fn((x.size(0), opt.nz))
            x_0 = generator(x, t, z)
            x = sample_posterior(coefficients, x_0, x, t, noise_fn(tuple(x.shape)))
    return x


class GraphSampler:
    """The whole T-step reverse loop (hot loop #1) captured once as a single CUDA graph around a
    GeneratorEngine: per step  z -> G(x, t, z) -> posterior update, all buffers static.

    Noise comes from static buffers `z_noise[T]`, `p_noise[T]`: parity runs fill them with the oracle's draws; throughput
    runs refill them with torch.randn (two launches per sampling call, outside the graph)."""

    def __init__(self, engine, args):
        self.eng = engine
        self.T = args.num_timesteps
        dev = engine.dev
        N = engine.N
        self.pc = PosteriorCoefficients(args, dev)
        self.xbuf = [torch.zeros_like(engine.x_in) for _ in range(2)]  # ping-pong: the update never aliases its input
        self.x = self.xbuf[0]
        self.z_noise = torch.zeros(self.T, N, engine.cfg.nz, device=dev)
        self.p_noise = torch.zeros(self.T, *engine.x_in.shape, device=dev)
        self.t_tab = torch.stack([torch.full((N,), i, dtype=torch.int64, device=dev) for i in range(self.T)])
        self.graph = None

    def _loop(self):
        e = self.eng
        cur = 0
        for k, i in enumerate(reversed(range(self.T))):
            e.x_in.copy_(self.xbuf[cur])
            e.t_in.copy_(self.t_tab[i])
            e.z_in.copy_(self.z_noise[k])
            e.run_steps()
            sample_posterior(self.pc, e.out, self.xbuf[cur], self.t_tab[i], self.p_noise[k], out=self.xbuf[1 - cur])
            cur = 1 - cur
        self.x = self.xbuf[cur]

    def capture(self):
        self._loop()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._loop()
        self.graph = g

    def sample(self, x_init, fresh_noise=True):
        self.xbuf[0].copy_(x_init)
        if fresh_noise:
            self.z_noise.normal_()
            self.p_noise.normal_()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._loop()
        return self.x
