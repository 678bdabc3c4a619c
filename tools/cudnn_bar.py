"""Kernel-level bar: what the reference's GPU path would run for its convolutions (nn.Conv2d -> cuDNN, NCHW fp32) on this
GPU, next to this repo's tcgen05 kernel on the same shapes.  fp32 without TF32 is the precision the parity gate asks for;
TF32 (PyTorch's default for convolutions) fails the 1e-4 gate (2.9e-4 per conv) and is listed for orientation."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
import torch.nn.functional as F
from ddgan_b200 import ops
dev = 'cuda'

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

print(f'{"shape":34s} {"cudnn fp32":>12s} {"cudnn tf32":>12s} {"conv_tc x3":>12s} {"conv_tc bf16":>13s}   (TFLOP/s algorithmic; us)')
for n, cin, cout, h, k in [(64, 128, 128, 32, 3), (64, 256, 256, 32, 3), (64, 256, 256, 16, 3), (64, 256, 256, 8, 3), (64, 256, 256, 4, 3),
                           (64, 256, 768, 16, 1)]:
    x = torch.randn(n, cin, h, h, device=dev); w = torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k); b = torch.zeros(cout, device=dev)
    fl = 2 * n * h * h * cout * cin * k * k
    res = []
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ms = timeit(lambda: F.conv2d(x, w, b, padding=k // 2))
        res.append(ms)
    mine = []
    for prec in (3, 1):
        xp = ops.to_pnhwc(x)
        taps = ops.TAPS_3X3 if k == 3 else ops.TAPS_1X1
        cw = ops.ConvWeights(cout, [(cin, len(taps))], dev, precision=prec, m_rows=n * (h + 2) * (h + 2) if k == 3 else n * h * h)
        cw.pack_conv_weight(0, w)
        out = ops.alloc_pnhwc(n, h, h, cout, dev)
        d = ops.build_conv_desc(cw, [ops.conv_src(xp, cin, taps)], n, h, h, out, bias=b)
        mine.append(timeit(lambda: ops.conv_launch(d)))
    f = lambda ms: f'{fl / ms / 1e9:6.1f} {ms * 1e3:5.0f}'
    print(f'{cin:4d}->{cout:<4d} {k}x{k} @{h:3d}px n={n:<3d}        {f(res[0]):>12s} {f(res[1]):>12s} {f(mine[0]):>12s} {f(mine[1]):>13s}')
