"""Repeat the SAME train step from the SAME state many times on one Trainer and compare every gradient tensor with the first run:
localises run-to-run nondeterminism (atomics are expected at 1e-7..1e-6; anything larger points at a race)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from oracle import ddgan_oracle as O
from ddgan_b200.modules import NCSNpp, Discriminator_small
from ddgan_b200.train import Trainer
DEV = 'cuda'
which = sys.argv[1] if len(sys.argv) > 1 else 'tiny'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
gs = int(sys.argv[3]) if len(sys.argv) > 3 else 0

def seeded(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))

if which == 'tiny':
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16); B = 4
    dshape = O.discriminator_param_shapes(6, 16, 32); ngf, ted = 16, 32
else:
    cfg = O.cifar10_config(); B = 8
    dshape = O.discriminator_param_shapes(6, 64, 256); ngf, ted = 64, 256
netG = NCSNpp(cfg).to(DEV); netG.load_state_dict(O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21))
netD = Discriminator_small(nc=6, ngf=ngf, t_emb_dim=ted).to(DEV); netD.load_state_dict(O.randomize_params(dshape, seed=22))
for k, v in dict(lr_g=1.6e-4, lr_d=1.25e-4, beta1_g=0.5, beta2_g=0.9, beta1_d=0.5, beta2_d=0.9, r1_gamma=0.02, lazy_reg=2, grad_clip_norm=1.0, ema_decay=0.999, use_ema=True).items():
    setattr(cfg, k, v)
tr = Trainer(cfg, netG, netD, DEV)
real = torch.tanh(seeded((B, 3, 32, 32), 300)).to(DEV)
nz = {}
for sfx, b in (('_d', 600), ('_g', 650)):
    nz['t' + sfx] = (torch.arange(B) % cfg.num_timesteps).to(DEV)
    for j, k in enumerate(('n_xtp1', 'n_xt', 'n_post')):
        nz[k + sfx] = seeded((B, 3, 32, 32), b + 1 + j).to(DEV)
    nz['z' + sfx] = seeded((B, cfg.nz), b + 9).to(DEV)
tr.step(real, 1, noise=nz)            # records + freezes the pack plans
snap = tr._snapshot()
ref = None
worst = {}
for r in range(reps):
    with torch.no_grad():
        for dst, src in snap:
            dst.copy_(src)
    e = tr.step(real, gs, noise=nz)
    cur = {'D': tr.optD.flat_g.clone(), 'G': tr.optG.flat_g.clone(), 'loss': (float(e[0]), float(e[1]))}
    if ref is None:
        ref = cur
        continue
    for tag, opt in (('D', tr.optD), ('G', tr.optG)):
        tot = O.rel_l2(cur[tag], ref[tag])
        line = f'rep {r:2d} {tag} total {tot:.2e}'
        big = []
        for nme, (off, num) in zip(opt.names, opt.views):
            a, b_ = ref[tag][off:off + num], cur[tag][off:off + num]
            d = float((a - b_).norm()) / max(float(a.norm()), 1e-30)
            worst[(tag, nme)] = max(worst.get((tag, nme), 0.0), d)
            if d > 1e-4 and float(a.norm()) > 1e-6 * float(ref[tag].norm()):
                big.append((d, nme))
        if tot > 2e-5 or big:
            print(line, ' losses', cur['loss'], ref['loss'])
            for d, nme in sorted(big, reverse=True)[:8]:
                print(f'      {nme:45s} rel diff {d:.2e}')
print('--- worst per-tensor relative difference over all repetitions (top 12) ---')
for (tag, nme), d in sorted(worst.items(), key=lambda kv: -kv[1])[:12]:
    print(f'{tag} {nme:45s} {d:.2e}')
