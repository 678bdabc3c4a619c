"""Per-shape timing of the tcgen05 weight-gradient kernel (BASELINE train step shapes)."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops
dev = 'cuda'
def run(n, cin, cout, h, k):
    cp = ops.pad_c(cin); cop = ops.pad_c(cout)
    x = ops.to_pnhwc(torch.randn(n, cin, h, h, device=dev), cpad=cp)
    dy = ops.to_pnhwc(torch.randn(n, cout, h, h, device=dev), cpad=cop)
    taps = ops.TAPS_3X3 if k == 3 else ops.TAPS_1X1
    dw = torch.zeros(cout, cin, k, k, device=dev)
    prof = torch.zeros(8, dtype=torch.int64, device=dev)
    f = lambda: ops.conv_wgrad(x, dy, dw, n, h + 2, h + 2, cout, cin, cp, taps, cin * k * k, k * k, 1, prof=prof)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2 * n * h * h * cout * cp * k * k
    print(f'wgrad {cin}->{cout} {k}x{k} @{h}px n={n}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s')
    pv = prof.cpu().tolist()
    print('   ' + '  '.join(f'{a}={b}' for a, b in zip(['prod_total', 'prod_wait_empty', 'wait_last_mma', 'staging', 'tiles', 'mma_total', 'mma_wait_full'], pv)))
for args in [(64, 128, 128, 32, 3), (64, 256, 128, 32, 3), (64, 256, 256, 32, 3), (64, 256, 256, 16, 3), (64, 512, 256, 16, 3), (64, 256, 256, 8, 3), (64, 256, 256, 4, 3),
             (64, 128, 128, 32, 1), (64, 256, 256, 16, 1), (64, 3, 128, 32, 3), (64, 128, 3, 32, 3), (64, 6, 64, 32, 3), (64, 64, 128, 32, 3), (64, 128, 256, 16, 3), (64, 256, 512, 8, 3), (64, 512, 512, 4, 3)]:
    run(*args)
