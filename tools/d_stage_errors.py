"""Per-stage relative-L2 error of the fused Discriminator_small engine at the BASELINE configuration (ngf 64, t_emb 256, B 64)
against the CPU oracle in float32 and in float64 (ground truth): shows where the logit error comes from."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
import torch.nn.functional as F
from oracle import ddgan_oracle as O
from ddgan_b200.engine import DiscriminatorEngine
from ddgan_b200 import ops

def seeded(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))

B = 64
sd = O.randomize_params(O.discriminator_param_shapes(6, 64, 256), seed=41)
x = seeded((B, 3, 32, 32), 420); xt = seeded((B, 3, 32, 32), 421); t = torch.arange(B) % 4
s32, s64 = [], []
with torch.no_grad():
    r32 = O.discriminator_forward(sd, x, t, xt, 256, stages=s32)
    r64 = O.discriminator_forward({k: v.double() for k, v in sd.items()}, x.double(), t, xt.double(), 256, stages=s64)
eng = DiscriminatorEngine(6, 64, 256, 32, B, large=False, device='cuda')
eng.load_state_dict(sd)
y = eng.forward(x.cuda(), t.cuda(), xt.cuda()).cpu()
names = ['start_conv'] + [f'conv{i + 1}' for i in range(4)]
for i, nme in enumerate(names):
    a = eng.stage_acts[i]
    g = ops.from_pnhwc(a.buf, s32[i].shape[1]).cpu()
    print(f'{nme:12s} gpu-vs-f64 {O.rel_l2(g, s64[i]):.3e}   cpu32-vs-f64 {O.rel_l2(s32[i], s64[i]):.3e}   gpu-vs-cpu32 {O.rel_l2(g, s32[i]):.3e}')
f = F.leaky_relu(ops.from_pnhwc(eng.final_feat.buf, s32[5].shape[1]).cpu(), 0.2)
print(f'final_conv   gpu-vs-f64 {O.rel_l2(f, s64[5]):.3e}   cpu32-vs-f64 {O.rel_l2(s32[5], s64[5]):.3e}   gpu-vs-cpu32 {O.rel_l2(f, s32[5]):.3e}')
p = eng.pooled.cpu()
print(f'pooled       gpu-vs-f64 {O.rel_l2(p, s64[6]):.3e}   cpu32-vs-f64 {O.rel_l2(s32[6], s64[6]):.3e}   gpu-vs-cpu32 {O.rel_l2(p, s32[6]):.3e}')
print(f'logit        gpu-vs-f64 {O.rel_l2(y, r64):.3e}   cpu32-vs-f64 {O.rel_l2(r32, r64):.3e}   gpu-vs-cpu32 {O.rel_l2(y, r32):.3e}')
print('logit rms', float(r64.pow(2).mean().sqrt()), 'pooled rms', float(s64[6].pow(2).mean().sqrt()))
