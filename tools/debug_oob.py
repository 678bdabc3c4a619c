import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch, torch.nn.functional as F
from ddgan_b200 import ops
from oracle import ddgan_oracle as O
dev = 'cuda'
torch.manual_seed(0)
N, C, H = 3, 64, 4
x = torch.randn(N, C, H, H)
xp = ops.to_pnhwc(x.to(dev))
sc = (torch.rand(N, C) + 0.5).to(dev); sh = (torch.randn(N, C) * 0.3).to(dev)
numel = N * 10 * 10 * C
guard = torch.full((3 * numel,), 7.0, device=dev)
out = guard[numel:2 * numel].view(N, 10, 10, C); out.zero_()
ops.fir_pnhwc(xp, 1, out, sc, sh, ops.ACT_SILU)
torch.cuda.synchronize()
print('fir guard before ok:', bool((guard[:numel] == 7).all()), 'after ok:', bool((guard[2 * numel:] == 7).all()))
ref = O.upsample_2d(F.silu(x * sc.cpu()[:, :, None, None] + sh.cpu()[:, :, None, None]))
print('fir err', O.rel_l2(ops.from_pnhwc(out.contiguous()).cpu(), ref))
# conv 64->64 3x3 at 8x8, N=3, no affine, with addvec and stats
h = torch.randn(N, C, 8, 8)
hp = ops.to_pnhwc(h.to(dev))
w = torch.randn(64, 64, 3, 3) / 24; b = torch.randn(64) * 0.1; tv = torch.randn(N, 200)
cw = ops.ConvWeights(64, [(64, 9)], dev); cw.pack_conv_weight(0, w.to(dev))
guard = torch.full((3 * numel,), 7.0, device=dev)
out = guard[numel:2 * numel].view(N, 10, 10, C); out.zero_()
st = torch.zeros(N * 64 * 2 + 64, dtype=torch.float64, device=dev); st[N * 64 * 2:] = 7
tvd = tv.to(dev)
ops.conv2d_fused(cw, [ops.conv_src(hp, 64, ops.TAPS_3X3)], N, 8, 8, out, bias=b.to(dev), stats=st, addvec=tvd.data_ptr() + 4 * 100, addvec_stride=200)
torch.cuda.synchronize()
print('conv guard before ok:', bool((guard[:numel] == 7).all()), 'after ok:', bool((guard[2 * numel:] == 7).all()), 'stats guard', bool((st[N * 64 * 2:] == 7).all()))
ref = F.conv2d(h, w, b, padding=1) + tv[:, 100:164, None, None]
y = ops.from_pnhwc(out.contiguous())
print('conv finite', bool(torch.isfinite(y).all()), 'err', O.rel_l2(y.cpu(), ref))
bad = ~torch.isfinite(y)
if bad.any():
    idx = bad.nonzero()
    print('bad count', int(bad.sum()), 'first', idx[:10].tolist())
