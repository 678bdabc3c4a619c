import os, sys, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
import bench
from ddgan_b200.modules import NCSNpp, Discriminator_small
from ddgan_b200.train import Trainer
B = int(sys.argv[1]); variants = tuple(sys.argv[2].split(','))
dev = torch.device('cuda')
cfg = bench.train_args()
netG = NCSNpp(cfg).to(dev); netD = Discriminator_small(nc=6, ngf=64, t_emb_dim=256).to(dev)
tr = Trainer(cfg, netG, netD, dev)
print('capturing', B, variants, flush=True)
tr.capture((B, 3, 32, 32), variants=variants)
torch.cuda.synchronize()
print('captured; mem GB', torch.cuda.memory_reserved() / 1e9, flush=True)
real = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
for i in range(3):
    gs = 1 if variants == ('plain',) else (0 if variants == ('r1',) else i * 15 % 16)
    out = tr.step_graphed(real, gs); torch.cuda.synchronize()
    print('replay', i, float(out[0]), float(out[1]), flush=True)
