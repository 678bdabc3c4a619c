"""One conv_tc launch per shape for ncu captures (tools/profile.sh style).  Usage: python tools/conv_one.py <h> [cin cout prec]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops
h = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cout = int(sys.argv[3]) if len(sys.argv) > 3 else 256
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n, dev = 64, 'cuda'
TAPS9 = [(dr, ds) for dr in (-1, 0, 1) for ds in (-1, 0, 1)]
x = ops.alloc_pnhwc(n, h, h, cin, dev); x[:, 1:-1, 1:-1, :] = torch.randn(n, h, h, cin, device=dev)
w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
cw = ops.ConvWeights(cout, [(cin, 9)], dev, precision=prec, m_rows=n * (h + 2) * (h + 2)); cw.pack_conv_weight(0, w)
scale = torch.rand(n, cin, device=dev) + 0.5; shift = torch.randn(n, cin, device=dev) * 0.1
out = ops.alloc_pnhwc(n, h, h, cout, dev)
st = torch.zeros(n, cout, 2, dtype=torch.float64, device=dev)
bias = torch.zeros(cout, device=dev)
d = ops.build_conv_desc(cw, [ops.conv_src(x, cin, TAPS9, scale=scale, shift=shift, act=ops.ACT_SILU)], n, h, h, out, stats=st, bias=bias, res=out)
for _ in range(3):
    ops.conv_launch(d)
torch.cuda.synchronize()
print(ops.conv_last_launch_info())
