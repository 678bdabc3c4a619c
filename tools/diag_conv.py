"""GPU bring-up diagnostic for the tcgen05 conv kernel: prints relative-L2 errors for a matrix of cases."""
import os, sys, math, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
import torch.nn.functional as F
from ddgan_b200 import ops
from oracle.ddgan_oracle import rel_l2

dev = 'cuda'
torch.manual_seed(0)

def run(name, n, cin, cout, h, k, swap=0, msub=0, prec=3, affine=False, act=0, res=False, nchw=False, stats=False):
    x = torch.randn(n, cin, h, h)
    w = torch.randn(cout, cin, k, k) / math.sqrt(cin * k * k)
    b = torch.randn(cout) * 0.1
    xin = x
    sc = sh = None
    if affine:
        sc = torch.rand(n, cin) + 0.5; sh = torch.randn(n, cin) * 0.3
        xin = x * sc[:, :, None, None] + sh[:, :, None, None]
    if act == 1: xin = F.silu(xin)
    if act == 2: xin = F.leaky_relu(xin, 0.2)
    ref = F.conv2d(xin.double(), w.double(), b.double(), padding=k // 2).float()
    r = None
    if res:
        r = torch.randn(n, cout, h, h)
        ref = (ref + r) / math.sqrt(2)
    cp = ops.pad_c(cin)
    xd = ops.to_pnhwc(x.to(dev), cpad=cp)
    taps = ops.TAPS_3X3 if k == 3 else ops.TAPS_1X1
    cw = ops.ConvWeights(cout, [(cp, len(taps))], dev, precision=prec, m_rows=n * (h + 2) * (h + 2) if k == 3 else n * h * h)
    cw.pack_conv_weight(0, w.to(dev).contiguous())
    scd = shd = None
    if affine:
        scd = torch.zeros(n, cp, device=dev); scd[:, :cin] = sc.to(dev)
        shd = torch.zeros(n, cp, device=dev); shd[:, :cin] = sh.to(dev)
    st = torch.zeros(n, cout, 2, dtype=torch.float64, device=dev) if stats else None
    if nchw:
        out = torch.zeros(n, cout, h, h, device=dev)
        mode = ops.OUT_NCHW
    else:
        out = ops.alloc_pnhwc(n, h, h, cout, dev)
        mode = ops.OUT_PNHWC
    rd = None
    if res:
        rd = r.to(dev).contiguous() if nchw else ops.to_pnhwc(r.to(dev), cpad=cout)
    t0 = time.time()
    ops.conv2d_fused(cw, [ops.conv_src(xd, cp, taps, scale=scd, shift=shd, act=act)], n, h, h, out, out_mode=mode,
                     bias=b.to(dev), res=rd, out_scale=(1 / math.sqrt(2) if res else 1.0), stats=st, msub=msub)
    torch.cuda.synchronize()
    y = out if nchw else ops.from_pnhwc(out, cout)
    err = rel_l2(y.cpu(), ref)
    extra = ''
    if stats:
        s1 = ref.double().sum(dim=(2, 3)); s2 = (ref.double() ** 2).sum(dim=(2, 3))
        e1 = rel_l2(st[:, :, 0].cpu(), s1); e2 = rel_l2(st[:, :, 1].cpu(), s2)
        extra = f' stats_err=({e1:.2e},{e2:.2e})'
    if not nchw:
        border = out.clone(); border[:, 1:-1, 1:-1, :] = 0
        extra += f' border_max={float(border.abs().max()):.1e}'
    print(f'{name:34s} swap={swap} msub={msub} prec={prec} rel_l2={err:.3e}{extra} ({(time.time()-t0)*1e3:.1f} ms)', flush=True)
    return err

if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    run('1x1 32->128 8px', 2, 32, 128, 8, 1, msub=1)
    run('3x3 32->128 8px', 2, 32, 128, 8, 3, msub=1)
    run('3x3 64->128 16px msub2', 4, 64, 128, 16, 3, msub=2)
    run('3x3 128->256 32px auto', 8, 128, 256, 32, 3)
    run('3x3 128->256 32px bf16', 8, 128, 256, 32, 3, prec=1)
    run('3x3 256->256 16px affine silu res', 4, 256, 256, 16, 3, affine=True, act=1, res=True, stats=True)
    run('3x3 3->128 32px (padded cin)', 4, 3, 128, 32, 3, stats=True)
    run('3x3 128->3 32px nchw', 4, 128, 3, 32, 3, nchw=True)
    run('1x1 128->64 4px leaky', 4, 128, 64, 4, 1, act=2)
    run('3x3 512->512 4px', 4, 512, 512, 4, 3)
