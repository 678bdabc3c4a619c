"""How far apart are two runs of the same train step from identical state?  eager vs eager (inherent: atomics order + gate flips)
and eager vs CUDA-graph replay, per iteration, with the per-tensor breakdown of the worst gradient difference."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from oracle import ddgan_oracle as O
from ddgan_b200.modules import NCSNpp, Discriminator_small
from ddgan_b200.train import Trainer
DEV = 'cuda'

def seeded(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))

def nets():
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
    netG = NCSNpp(cfg).to(DEV); netG.load_state_dict(O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21))
    netD = Discriminator_small(nc=6, ngf=16, t_emb_dim=32).to(DEV); netD.load_state_dict(O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22))
    for k, v in dict(lr_g=1.6e-4, lr_d=1.25e-4, beta1_g=0.5, beta2_g=0.9, beta1_d=0.5, beta2_d=0.9, r1_gamma=0.02, lazy_reg=2, grad_clip_norm=1.0, ema_decay=0.999, use_ema=True).items():
        setattr(cfg, k, v)
    return cfg, netG, netD

def noise(cfg, B, S, base):
    nz = {}
    for sfx, b in (('_d', base), ('_g', base + 50)):
        nz['t' + sfx] = (torch.arange(B) % cfg.num_timesteps).to(DEV)
        for j, k in enumerate(('n_xtp1', 'n_xt', 'n_post')):
            nz[k + sfx] = seeded((B, 3, S, S), b + 1 + j).to(DEV)
        nz['z' + sfx] = seeded((B, cfg.nz), b + 9).to(DEV)
    return nz

def sync(a, b):
    for oa, ob in ((a.optD, b.optD), (a.optG, b.optG)):
        for ta, tb in ((oa.flat_p, ob.flat_p), (oa.m, ob.m), (oa.v, ob.v), (oa.state, ob.state)):
            tb.copy_(ta)
    b.optG.ema.copy_(a.optG.ema)

B = 4
real = torch.tanh(seeded((B, 3, 32, 32), 300)).to(DEV)
for mode in ('eager-eager', 'eager-graph'):
    cfg, netG, netD = nets()
    netG2, netD2 = copy.deepcopy(netG), copy.deepcopy(netD)
    a = Trainer(cfg, netG, netD, DEV); b = Trainer(cfg, netG2, netD2, DEV)
    if mode == 'eager-graph':
        b.capture((B, 3, 32, 32), warmup=3)
    for it in range(8):
        nz = noise(cfg, B, 32, 600 + 10 * it)
        ea = a.step(real, it, noise=nz)
        eb = (b.step_graphed if mode == 'eager-graph' else b.step)(real, it, noise=nz)
        eD = O.rel_l2(b.optD.flat_g.cpu(), a.optD.flat_g.cpu()); eG = O.rel_l2(b.optG.flat_g.cpu(), a.optG.flat_g.cpu())
        print(f'{mode} it={it} r1={it % 2 == 0}  dLossD {abs(float(ea[0]) - float(eb[0])):.2e} dLossG {abs(float(ea[1]) - float(eb[1])):.2e}  gradD {eD:.2e} gradG {eG:.2e}')
        if eG > 1e-4 or eD > 1e-4:
            for opt_a, opt_b, tag in ((a.optD, b.optD, 'D'), (a.optG, b.optG, 'G')):
                rows = []
                for nme, (off, num) in zip(opt_a.names, opt_a.views):
                    ga, gb = opt_a.flat_g[off:off + num], opt_b.flat_g[off:off + num]
                    rows.append((float((ga - gb).norm()), float(ga.norm()), nme))
                rows.sort(reverse=True)
                tot = float(opt_a.flat_g.norm())
                print(f'   {tag}: |g| = {tot:.3e}; largest absolute differences:')
                for d, n_, nme in rows[:6]:
                    print(f'      {nme:45s} |diff| {d:.3e}  |g| {n_:.3e}')
        sync(a, b)
