import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops
dev = 'cuda'
names = ['prod_total', 'prod_wait_emptyA', 'prod_wait_acc', 'epilogue', 'loader_total', 'loader_wait_emptyB', 'mma_total', 'mma_wait_A', 'mma_wait_B', 'mma_wait_accEmpty', 'tiles', 'epi_total', 'epi_wait_accFull']
def run(n, cin, cout, h, k, msub=0, affine=True, prec=3, nt256=1, stats=True):
    from ddgan_b200._lib import lib
    lib().ddg_conv_set_nt256(nt256)
    cp = ops.pad_c(cin)
    x = ops.to_pnhwc(torch.randn(n, cin, h, h, device=dev), cpad=cp)
    w = torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k)
    taps = ops.TAPS_3X3 if k == 3 else ops.TAPS_1X1
    cw = ops.ConvWeights(cout, [(cp, len(taps))], dev, precision=prec, m_rows=n * (h + 2) * (h + 2) if k == 3 else n * h * h); cw.pack_conv_weight(0, w)
    sc = torch.rand(n, cp, device=dev) + 0.5 if affine else None; sh = torch.randn(n, cp, device=dev) if affine else None
    out = ops.alloc_pnhwc(n, h, h, cout, dev)
    st = torch.zeros(n, cout, 2, dtype=torch.float64, device=dev)
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    d = ops.build_conv_desc(cw, [ops.conv_src(x, cp, taps, scale=sc, shift=sh, act=1 if affine else 0)], n, h, h, out, stats=st if stats else None, msub=msub, prof=prof,
                            bias=torch.zeros(cout, device=dev))
    for _ in range(3): ops.conv_launch(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.conv_launch(d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2 * n * h * h * cout * cp * k * k
    pv = prof.cpu().tolist()
    print(f'conv {cin}->{cout} {k}x{k} @{h}px n={n} msub={msub} prec={prec} nt256={nt256} stats={stats}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s')
    print('   ' + '  '.join(f'{a}={b}' for a, b in zip(names, pv)))
if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'big'
    if which == 'big':
        run(64, 128, 128, 32, 3)
        run(64, 256, 256, 32, 3)
        run(64, 256, 768, 16, 1)
    else:   # the small levels (4x4, 8x8): N = 64 / 128 one-tile CTAs
        run(64, 256, 256, 4, 3)
        run(64, 256, 256, 8, 3)
        run(64, 256, 256, 8, 3, nt256=1 | 2)
        run(64, 256, 256, 4, 3, affine=False, stats=False)
        run(64, 256, 256, 16, 3)
