"""Bisect which part of the train step breaks CUDA-graph replay: each stage runs in its own subprocess."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'denoising-diffusion-gan_b200'))
import torch, torch.nn.functional as F
from oracle import ddgan_oracle as O
from ddgan_b200.modules import NCSNpp, Discriminator_small
from ddgan_b200 import diffusion, ops, train_graph as TG
stage = sys.argv[1]
dev = torch.device('cuda')
cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
netG = NCSNpp(cfg).to(dev); netD = Discriminator_small(nc=6, ngf=16, t_emb_dim=32).to(dev)
x = torch.randn(4, 3, 32, 32, device=dev); xt = torch.randn(4, 3, 32, 32, device=dev); t = torch.tensor([0, 1, 2, 3], device=dev)
z = torch.randn(4, cfg.nz, device=dev)
optD = torch.optim.Adam(netD.parameters(), lr=1e-4, capturable=True, foreach=True)
def body():
    if stage == 'conv_fwd':
        xp = TG.ToPnhwcFn.apply(x, 32)
        return TG.conv3x3(xp, netG.all_modules[2].weight, netG.all_modules[2].bias, 4, 32, 32).sum()
    if stage == 'conv_bwd':
        netG.zero_grad(set_to_none=True)
        xp = TG.ToPnhwcFn.apply(x, 32)
        y = TG.conv3x3(xp, netG.all_modules[2].weight, netG.all_modules[2].bias, 4, 32, 32).sum(); y.backward(); return y
    if stage == 'gn_bwd':
        xr = x.clone().requires_grad_(True)
        y = TG.group_norm_act(TG.ToPnhwcFn.apply(xr, 32), 32, 32, 8, torch.ones(32, device=dev), torch.zeros(32, device=dev), 1).sum(); y.backward(); return y
    if stage == 'fir_bwd':
        xr = x.clone().requires_grad_(True)
        y = TG.fir_down(TG.fir_up(TG.ToPnhwcFn.apply(xr, 32))).sum(); y.backward(); return y
    if stage == 'lin_bwd':
        netG.zero_grad(set_to_none=True)
        y = TG.LinearFn.apply(z, netG.z_transform[1].weight, netG.z_transform[1].bias).sum(); y.backward(); return y
    if stage == 'd_fwd':
        with torch.no_grad():
            return TG.discriminator_forward(netD, x, t, xt).sum()
    if stage == 'd_bwd':
        netD.zero_grad(set_to_none=True)
        y = F.softplus(-netD(x, t, xt)).mean(); y.backward(); return y
    if stage == 'd_r1':
        netD.zero_grad(set_to_none=True)
        xr = x.clone().requires_grad_(True)
        d = netD(xr, t, xt).view(-1)
        g = torch.autograd.grad(d.sum(), xr, create_graph=True)[0]
        gp = (g.view(4, -1).norm(2, dim=1) ** 2).mean(); gp.backward(); return gp
    if stage == 'd_opt':
        netD.zero_grad(set_to_none=True)
        y = F.softplus(-netD(x, t, xt)).mean(); y.backward()
        torch.nn.utils.clip_grad_norm_(netD.parameters(), 1.0); optD.step(); return y
    if stage == 'g_bwd':
        netG.zero_grad(set_to_none=True)
        y = netG(x, t, z).sum(); y.backward(); return y
    if stage == 'rand':
        return torch.randn(4, 8, device=dev).sum() + torch.randint(0, 4, (4,), device=dev).sum()
if stage.startswith('tr_'):
    from ddgan_b200.train import Trainer
    for k, v in dict(lr_g=1.6e-4, lr_d=1.25e-4, beta1=0.5, beta2=0.9, r1_gamma=0.02, lazy_reg=2, grad_clip_norm=1.0, ema_decay=0.999, use_ema=(stage != 'tr_noema')).items():
        setattr(cfg, k, v)
    tr = Trainer(cfg, netG, netD, dev)
    variants = {'tr_plain': ('plain',), 'tr_r1': ('r1',)}.get(stage, ('r1', 'plain'))
    tr.capture((4, 3, 32, 32), variants=variants, share_pool=(stage == 'tr_shared'))
    print(stage, 'captured', flush=True)
    for i in range(4):
        gs = 1 if stage == 'tr_plain' else (0 if stage == 'tr_r1' else i)
        out = tr.step_graphed(x, gs)
        torch.cuda.synchronize()
    print(stage, 'replayed', float(out[0]), float(out[1]), flush=True)
    sys.exit(0)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): body()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = body()
torch.cuda.synchronize(); print(stage, 'captured', flush=True)
g.replay(); torch.cuda.synchronize(); print(stage, 'replayed', float(out), flush=True)
''' % (ROOT, ROOT)
open('/tmp/stage.py', 'w').write(STAGE)
for st in sys.argv[1:]:
    r = subprocess.run([sys.executable, '/tmp/stage.py', st], capture_output=True, text=True, timeout=200)
    tail = (r.stdout.strip().splitlines() or [''])[-1]
    err = [l for l in r.stderr.splitlines() if 'Error' in l or 'error' in l][:2]
    print(f'{st:10s} rc={r.returncode} {tail} {err}', flush=True)
