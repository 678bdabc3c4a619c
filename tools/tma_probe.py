"""Per-layer A/B of the three ways conv_tc can be fed: fp32 source with the GN+SiLU prologue, fp32 source without a prologue, and
pre-split bf16 planes fetched by TMA.  Prints microseconds per launch and TFLOP/s (algorithmic) for both precisions.
Usage: python tools/tma_probe.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops

TAPS9 = [(dr, ds) for dr in (-1, 0, 1) for ds in (-1, 0, 1)]


def time_fn(fn, iters=30):
    for _ in range(5):
        fn()
    flush = torch.empty(160 << 20, dtype=torch.uint8, device='cuda')
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3


def run(n, h, cin, cout, prec, nsrc=1):
    dev = 'cuda'
    xs = []
    for _ in range(nsrc):
        x = ops.alloc_pnhwc(n, h, h, cin, dev)
        x[:, 1:-1, 1:-1, :] = torch.randn(n, h, h, cin, device=dev)
        xs.append(x)
    w = torch.randn(cout, cin * nsrc, 3, 3, device=dev) * 0.05
    cw = ops.ConvWeights(cout, [(cin, 9)] * nsrc, dev, precision=prec, m_rows=n * h * h)
    for i in range(nsrc):
        cw.pack_conv_weight(i, w[:, i * cin:(i + 1) * cin].contiguous())
    scale = torch.rand(n, cin, device=dev) + 0.5
    shift = torch.randn(n, cin, device=dev) * 0.1
    out = ops.alloc_pnhwc(n, h, h, cout, dev)
    planes = [ops.split_planes(x, ops.alloc_planes(n, h, h, cin, prec, dev), prec) for x in xs]
    flops = 2.0 * n * h * h * cout * cin * 9 * nsrc
    res = {}
    for name, mk in (('prologue', lambda i: ops.conv_src(xs[i], cin, TAPS9, scale=scale, shift=shift, act=ops.ACT_SILU)),
                     ('plain', lambda i: ops.conv_src(xs[i], cin, TAPS9)),
                     ('tma', lambda i: ops.conv_src(xs[i], cin, TAPS9, planes=planes[i]))):
        d = ops.build_conv_desc(cw, [mk(i) for i in range(nsrc)], n, h, h, out)
        us = time_fn(lambda: ops.conv_launch(d))
        res[name] = us
    info = ops.conv_last_launch_info()
    print(f'prec {prec} n{n} {h}x{h} cin {cin}x{nsrc} cout {cout} tile(msub,nt,persist,ctas)={info}: ' +
          '  '.join(f'{k} {v:7.1f} us {flops / v / 1e6:6.1f} TF/s' for k, v in res.items()), flush=True)


if __name__ == '__main__':
    for prec in (3, 1):
        run(64, 32, 128, 128, prec)
        run(64, 32, 128, 128, prec, nsrc=2)
        run(64, 32, 256, 256, prec)
        run(64, 16, 256, 256, prec)
        run(64, 16, 256, 256, prec, nsrc=2)
        run(64, 8, 256, 256, prec)
        run(64, 8, 256, 256, prec, nsrc=2)
        run(64, 4, 256, 256, prec)
        run(64, 4, 256, 256, prec, nsrc=2)
