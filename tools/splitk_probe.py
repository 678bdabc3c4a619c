"""A/B of the split-K path on the small spatial levels: microseconds per launch with and without the workspace.
Usage: python tools/splitk_probe.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops
from ddgan_b200._lib import lib
from tma_probe import time_fn, TAPS9


def run(n, h, cin, cout, prec, nsrc=1):
    dev = 'cuda'
    xs = []
    for _ in range(nsrc):
        x = ops.alloc_pnhwc(n, h, h, cin, dev)
        x[:, 1:-1, 1:-1, :] = torch.randn(n, h, h, cin, device=dev)
        xs.append(x)
    w = torch.randn(cout, cin * nsrc, 3, 3, device=dev) * 0.05
    cw = ops.ConvWeights(cout, [(cin, 9)] * nsrc, dev, precision=prec, m_rows=n * (h + 2) * (h + 2))
    for i in range(nsrc):
        cw.pack_conv_weight(i, w[:, i * cin:(i + 1) * cin].contiguous())
    scale = torch.rand(n, cin, device=dev) + 0.5
    shift = torch.randn(n, cin, device=dev) * 0.1
    out = ops.alloc_pnhwc(n, h, h, cout, dev)
    st = torch.zeros(n, cout, 2, dtype=torch.float64, device=dev)
    ws = ops.alloc_splitk_ws(dev)
    flops = 2.0 * n * h * h * cout * cin * 9 * nsrc
    res = {}
    for name, w_ in (('one CTA per tile', None), ('split-K', ws)):
        d = ops.build_conv_desc(cw, [ops.conv_src(xs[i], cin, TAPS9, scale=scale, shift=shift, act=ops.ACT_SILU) for i in range(nsrc)], n, h, h, out,
                                stats=st, splitk_ws=w_)
        us = time_fn(lambda: ops.conv_launch(d))
        res[name] = (us, lib().ddg_conv_last_launch_ksplit(), ops.conv_last_launch_info())
    print(f'prec {prec} n{n} {h}x{h} cin {cin}x{nsrc} cout {cout}: ' +
          '  '.join(f'{k}: {v[0]:6.1f} us {flops / v[0] / 1e6:6.1f} TF/s (ksplit {v[1]}, tile {v[2]})' for k, v in res.items()), flush=True)


if __name__ == '__main__':
    for prec in (3, 1):
        run(64, 4, 256, 256, prec)
        run(64, 4, 256, 256, prec, nsrc=2)
        run(64, 8, 256, 256, prec)
        run(16, 8, 256, 256, prec)
        run(16, 4, 512, 512, prec)
        run(8, 16, 256, 256, prec)
