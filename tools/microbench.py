"""BASELINE.json configs[4]: upfirdn2d / fused_leaky_relu / GroupNorm(+AdaGN+SiLU) microbench sweep vs the HBM roofline.
Algorithmic bytes per SURVEY.md 8(d): upfirdn2d 4*(in+out); fused_leaky_relu 8*numel; GroupNorm fwd 8*numel.
L2 is flushed between timed launches by READING a 512 MB buffer (clean lines: a memset would leave 126 MB of dirty lines whose
write-back is billed to the kernel under test)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import ops

dev = 'cuda'
peak = 6542.1
p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p))['hbm_gbs']
flush = torch.zeros(512 * 1024 * 1024 // 4, device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps

rows = []
k4 = torch.tensor([1., 3., 3., 1.]); k4 = torch.outer(k4, k4); k4 = (k4 / k4.sum()).to(dev)
N = 64
for C in (64, 128, 256, 512):
    for H in (32, 64, 128, 256):
        numel = N * C * H * H
        if numel * 4 * 5 > 12e9:
            continue
        x = torch.randn(N * C, H, H, device=dev)
        ms = timeit(lambda: ops.upfirdn2d_raw(x, k4, 1, 1, 2, 2, 1, 1, 1, 1))
        by = 4 * numel * 1.25
        rows.append(('upfirdn2d_down2', C, H, ms, by / ms / 1e6))
        if numel * 4 * 5 <= 8e9:
            ms = timeit(lambda: ops.upfirdn2d_raw(x, k4 * 4, 2, 2, 1, 1, 2, 1, 2, 1))
            rows.append(('upfirdn2d_up2', C, H, ms, 4 * numel * 5 / ms / 1e6))
        x4 = x.view(N, C, H, H)
        b = torch.randn(C, device=dev)
        ms = timeit(lambda: ops.fused_bias_act(x4, b, None, 3, 0, 0.2, 2 ** 0.5))
        rows.append(('fused_leaky_relu', C, H, ms, 8 * numel / ms / 1e6))
        G = min(C // 4, 32)
        gamma = torch.randn(N, C, device=dev); beta = torch.randn(N, C, device=dev)
        ms = timeit(lambda: ops.groupnorm_fwd(x4, G, gamma, beta, per_sample=True, act=ops.ACT_SILU))
        rows.append(('adagn_silu_fwd', C, H, ms, 8 * numel / ms / 1e6))
        del x, x4
print(f'{"op":18s} {"C":>4s} {"HW":>4s} {"ms":>9s} {"GB/s":>8s} {"frac":>6s}   (peak {peak:.0f} GB/s measured copy)')
for r in rows:
    print(f'{r[0]:18s} {r[1]:4d} {r[2]:4d} {r[3]:9.4f} {r[4]:8.0f} {r[4]/peak:6.2f}')
