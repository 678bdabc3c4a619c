"""Runs each non-conv hot-path operator a few times at a representative size (for `ncu --set full -k regex:...`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops
dev = 'cuda'
N = 64
k4 = torch.tensor([1., 3., 3., 1.]); k4 = torch.outer(k4, k4); k4 = (k4 / k4.sum()).to(dev)
x = torch.randn(N * 256, 64, 64, device=dev)
xg = torch.randn(N, 128, 64, 64, device=dev); gamma = torch.randn(N, 128, device=dev); beta = torch.randn(N, 128, device=dev)
xp = ops.to_pnhwc(torch.randn(N, 256, 16, 16, device=dev)); sc = torch.rand(N, 256, device=dev) + 0.5; sh = torch.randn(N, 256, device=dev)
up_out = ops.alloc_pnhwc(N, 32, 32, 256, dev)
xw = ops.to_pnhwc(torch.randn(N, 128, 32, 32, device=dev)); dyw = ops.to_pnhwc(torch.randn(N, 128, 32, 32, device=dev))
dw = torch.zeros(128, 128, 3, 3, device=dev)
b = torch.randn(256, device=dev); xb = torch.randn(N, 256, 64, 64, device=dev)
for _ in range(3):
    ops.upfirdn2d_raw(x, k4, 1, 1, 2, 2, 1, 1, 1, 1)                      # down x2, 256 ch, 64 px
    ops.upfirdn2d_raw(x, k4 * 4, 2, 2, 1, 1, 2, 1, 2, 1)                  # up x2
    ops.groupnorm_fwd(xg, 32, gamma, beta, per_sample=True, act=ops.ACT_SILU)
    ops.fused_bias_act(xb, b, None, 3, 0, 0.2, 2 ** 0.5)
    ops.fir_pnhwc(xp, 1, up_out, sc, sh, ops.ACT_SILU)                    # PNHWC up x2 with fused AdaGN + SiLU, 16 -> 32 px
    ops.conv_wgrad(xw, dyw, dw, N, 34, 34, 128, 128, 128, ops.TAPS_3X3, 128 * 9, 9, 1)
torch.cuda.synchronize(); print('ok')
