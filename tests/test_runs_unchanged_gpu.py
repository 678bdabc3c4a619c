""""train_ddgan.py / test_ddgan.py run unchanged": the reference's OWN scripts (staged, unmodified, under git-ignored baseline/_ref
by tools/stage_reference.py) imported on top of this repository's drop-in `score_sde` package:

  * test_ddgan.sample_from_model + test_ddgan.Posterior_Coefficients drive our NCSNpp, result checked against the CPU oracle fed
    with the very same random draws;
  * the statements of the reference training loop body (ddgan.py, `for iteration, (x, _) in enumerate(data_loader):` ... EMA step)
    are exec'ed verbatim from the staged source with our NCSNpp / Discriminator_small wrapped in DistributedDataParallel (world 1),
    the reference's own q_sample_pairs / sample_posterior / EMA and torch.optim.Adam; two iterations; the losses of both
    iterations are checked against a CPU replica of the same loop built from the oracle (iteration 2 sees the Adam updates of
    iteration 1, so the parameter updates are covered too).
Third-party imports of the scripts that have nothing to do with the path (datasets, PSO optimiser, FID, nibabel) are stubbed."""
import os
import sys
import textwrap
import types
from argparse import Namespace

import pytest
import torch
import torch.nn.functional as F

from oracle import ddgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, 'baseline', '_ref')


def _import_staged():
    if not os.path.exists(os.path.join(STAGED, 'ddgan.py')):
        pytest.skip('baseline/_ref not staged (tools/stage_reference.py runs in the build container)')
    for name, attrs in (('nibabel', ()), ('datasets_prep', ()), ('datasets_prep.custom', ('DatasetCustom', 'PositivePatchDataset', 'Luna16Dataset')),
                        ('pso_optim', ('AdaptivePSO',)), ('pytorch_fid', ()), ('pytorch_fid.fid_score', ('calculate_fid_given_paths',))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m
    import score_sde.models.ncsnpp_generator_adagn as ours
    assert 'denoising-diffusion-gan_b200' in ours.__file__            # the drop-in package, not the reference's
    if STAGED not in sys.path:
        sys.path.append(STAGED)
    import ddgan as ref_ddgan
    import test_ddgan as ref_test
    assert ref_ddgan.__file__.startswith(STAGED) and ref_test.__file__.startswith(STAGED)
    assert ref_ddgan.NCSNpp is ours.NCSNpp and ref_test.NCSNpp is ours.NCSNpp
    return ref_ddgan, ref_test


def _args(cfg, **kw):
    d = dict(vars(cfg))
    d.update(kw)
    return Namespace(**d)


def test_reference_sampling_script_drives_the_drop_in_generator():
    ref_ddgan, ref_test = _import_staged()
    cfg = O.tiny_config()
    args = _args(cfg)
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=7)
    netG = ref_test.NCSNpp(args).to(DEV)
    netG.load_state_dict(sd, strict=True)          # test_ddgan.py:162
    netG.eval()
    pos_coeff = ref_test.Posterior_Coefficients(args, DEV)
    T = ref_test.get_time_schedule(args, DEV)
    B = 3
    torch.manual_seed(1234)
    x_init = torch.randn(B, 3, cfg.image_size, cfg.image_size, device=DEV)
    y = ref_test.sample_from_model(pos_coeff, netG, args.num_timesteps, x_init, T, args)
    # replay the draws (test_ddgan.py:116-125: latent z, then the posterior noise, per step) for the oracle
    torch.manual_seed(1234)
    x0 = torch.randn(B, 3, cfg.image_size, cfg.image_size, device=DEV)
    draws = []
    for _ in range(args.num_timesteps):
        draws.append(torch.randn(B, cfg.nz, device=DEV).cpu())
        draws.append(torch.randn_like(x0).cpu())
    it = iter(draws)
    ref = O.sample_from_model(O.posterior_coefficients(cfg), lambda x, t, z: O.ncsnpp_forward(sd, cfg, x, t, z), cfg.num_timesteps,
                              x0.cpu(), cfg.nz, lambda shape: next(it))
    assert O.rel_l2(y.cpu(), ref) < 1e-4


def _loop_body_source(path):
    src = open(path).read().splitlines()
    a = next(i for i, l in enumerate(src) if l.strip().startswith('for iteration, (x, _) in enumerate(data_loader):'))
    b = next(i for i, l in enumerate(src) if i > a and l.strip().startswith("if args.kind_of_optim.lower() == 'adam' and not args.no_lr_decay"))
    return textwrap.dedent('\n'.join(src[a:b]))


def test_reference_training_loop_body_runs_verbatim_on_the_drop_in_modules():
    import torch.distributed as dist
    import torch.nn as nn
    ref_ddgan, _ = _import_staged()
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
    B, iters = 4, 2
    args = _args(cfg, kind_of_optim='adam', lazy_reg=1, r1_gamma=0.05, grad_clip_norm=1.0, use_ema=True, ema_decay=0.999,
                 lr_g=1.6e-4, lr_d=1.25e-4, beta1_g=0.5, beta2_g=0.9, beta1_d=0.5, beta2_d=0.9, weight_decay_G=0.0, weight_decay_D=0.0,
                 batch_size=B, no_lr_decay=True)
    sd_g = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21)
    sd_d = O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22)
    own_pg = not dist.is_initialized()
    if own_pg:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', str(29600 + os.getpid() % 300))
        dist.init_process_group('nccl', rank=0, world_size=1)
    try:
        torch.cuda.set_device(0)
        device = torch.device('cuda:0')
        # ---- ddgan.py:274-365 on the drop-in modules ----
        netG = ref_ddgan.NCSNpp(args).to(device)
        netG.load_state_dict(sd_g, strict=True)
        netD = ref_ddgan.Discriminator_small(nc=2 * args.num_channels, ngf=args.ngf, t_emb_dim=args.t_emb_dim, act=nn.LeakyReLU(0.2)).to(device)
        netD.load_state_dict(sd_d, strict=True)
        ref_ddgan.broadcast_params(netG.parameters())
        ref_ddgan.broadcast_params(netD.parameters())
        optimizerD = torch.optim.Adam(netD.parameters(), lr=args.lr_d, betas=(args.beta1_d, args.beta2_d), weight_decay=args.weight_decay_D)
        optimizerG = torch.optim.Adam(netG.parameters(), lr=args.lr_g, betas=(args.beta1_g, args.beta2_g), weight_decay=args.weight_decay_G)
        emaG = ref_ddgan.EMA(netG, optimizerG, ema_decay=args.ema_decay, device=device)
        netG = nn.parallel.DistributedDataParallel(netG, device_ids=[0])
        netD = nn.parallel.DistributedDataParallel(netD, device_ids=[0])
        coeff = ref_ddgan.DiffusionCoefficients(args, device)
        pos_coeff = ref_ddgan.PosteriorCoefficients(args, device)
        reals = [torch.tanh(torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(900 + i))) for i in range(iters)]
        ns = dict(vars(ref_ddgan))
        ns.update(args=args, netG=netG, netD=netD, optimizerD=optimizerD, optimizerG=optimizerG, emaG=emaG, coeff=coeff, pos_coeff=pos_coeff,
                  device=device, batch_size=B, nz=args.nz, rank=0, epoch=0, global_step=0, limited_iter=None,
                  data_loader=[(r, None) for r in reals], loss_values_D=[], loss_values_G=[], local_loss_D=[], local_loss_G=[])
        torch.manual_seed(4321)
        exec(compile(_loop_body_source(os.path.join(STAGED, 'ddgan.py')), 'ddgan.py[loop body]', 'exec'), ns)
        got_D, got_G = ns['loss_values_D'], ns['loss_values_G']
        assert ns['global_step'] == iters and len(got_D) == iters
        ema_state = {k: v.clone() for k, v in emaG.ema_state.items()}
        # ---- replay the CUDA draws of the loop (ddgan.py:450,122,112,470,164 then :491-497) ----
        torch.manual_seed(4321)
        shp = (B, 3, 32, 32)
        draws = []
        for _ in range(iters):
            d = {}
            for sfx in ('_d', '_g'):
                d['t' + sfx] = torch.randint(0, args.num_timesteps, (B,), device=device).cpu()
                d['n_xtp1' + sfx] = torch.randn(shp, device=device).cpu()
                d['n_xt' + sfx] = torch.randn(shp, device=device).cpu()
                d['z' + sfx] = torch.randn(B, args.nz, device=device).cpu()
                d['n_post' + sfx] = torch.randn(shp, device=device).cpu()
            draws.append(d)
    finally:
        if own_pg:
            dist.destroy_process_group()
    # ---- CPU replica from the oracle: same losses, clip, Adam ----
    pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    pd = {k: v.clone().requires_grad_(True) for k, v in sd_d.items()}
    oD = torch.optim.Adam(list(pd.values()), lr=args.lr_d, betas=(args.beta1_d, args.beta2_d))
    oG = torch.optim.Adam(list(pg.values()), lr=args.lr_g, betas=(args.beta1_g, args.beta2_g))
    ema = {k: v.detach().clone() for k, v in pg.items()}
    exp_D, exp_G = [], []
    for i in range(iters):
        d = draws[i]
        oD.zero_grad(set_to_none=True); oG.zero_grad(set_to_none=True)
        er, gp, ef = O.d_step_losses(pg, pd, cfg, reals[i], d['t_d'], (d['n_xt_d'], d['n_xtp1_d'], d['n_post_d']), d['z_d'], args.r1_gamma, do_r1=True)
        (er + gp + ef).backward()
        exp_D.append(float(er + ef))
        torch.nn.utils.clip_grad_norm_(list(pd.values()), args.grad_clip_norm)
        oD.step()
        oG.zero_grad(set_to_none=True)
        eg = O.g_step_loss(pg, {k: v.detach() for k, v in pd.items()}, cfg, reals[i], d['t_g'], (d['n_xt_g'], d['n_xtp1_g'], d['n_post_g']), d['z_g'])
        eg.backward()
        exp_G.append(float(eg))
        torch.nn.utils.clip_grad_norm_(list(pg.values()), args.grad_clip_norm)
        oG.step()
        for k in ema:
            ema[k].mul_(args.ema_decay).add_(pg[k].detach(), alpha=1 - args.ema_decay)
    for i in range(iters):
        # iteration 2 runs on weights updated by iteration 1 (sign-like first Adam steps): 1e-3 there, 1e-4 on the first
        tol = 1e-4 if i == 0 else 1e-3
        assert abs(got_D[i] - exp_D[i]) < tol * max(1.0, abs(exp_D[i])), (i, got_D[i], exp_D[i])
        assert abs(got_G[i] - exp_G[i]) < tol * max(1.0, abs(exp_G[i])), (i, got_G[i], exp_G[i])
    worst = max(O.rel_l2(ema_state[k].cpu(), ema[k]) for k in ema)
    assert worst < 1e-4, worst
