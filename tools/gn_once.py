import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops
N, C, H = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 64
x = torch.randn(N, C, H, H, device='cuda'); G = min(C // 4, 32)
gamma = torch.randn(N, C, device='cuda'); beta = torch.randn(N, C, device='cuda')
for _ in range(3):
    ops.groupnorm_fwd(x, G, gamma, beta, per_sample=True, act=ops.ACT_SILU)
torch.cuda.synchronize(); print('ok')
