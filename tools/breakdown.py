"""Per-step eager timing of one generator forward (CUDA events around every plan step)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import arch
from ddgan_b200.engine import GeneratorEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = arch.make_config()
if len(sys.argv) > 3 and sys.argv[3] == 'hq256':
    cfg = arch.make_config(image_size=256, num_channels_dae=64, ch_mult=(1, 1, 2, 2, 4, 4), n_mlp=3, num_timesteps=2, ngf=64)
eng = GeneratorEngine(cfg, B, 'cuda', precision=prec)
sd = {k: torch.randn(s) * (0.05 if len(s) > 1 else 0.1) + (1.0 if (len(s) == 1 and k.endswith('weight')) else 0.0) for k, s in eng.shapes.items()}
eng.load_state_dict(sd)
eng.x_in.normal_(); eng.z_in.normal_()
for _ in range(3):
    eng.run_steps()
torch.cuda.synchronize()
reps = 5
acc = [0.0] * len(eng.steps)
for r in range(reps):
    evs = []
    for st in eng.steps:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        acc[i] += a.elapsed_time(b) / reps
tot = sum(acc)
print(f'total {tot:.3f} ms over {len(acc)} steps')
flops = {}
for (cw, desc, srcs, kw, out), in [(k,) for k in eng._keep if isinstance(k, tuple) and len(k) == 5 and hasattr(k[1], 'nsrc')]:
    pass
agg = collections.OrderedDict()
for t, n in zip(acc, eng.step_names):
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += t
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
for n, (c, t) in rows[:45]:
    extra = ''
    if n.startswith('conv'):
        import re
        m = re.search(r'cout=(\d+) (\d+)x(\d+) srcs=(.*)', n)
        cout, h, w = int(m.group(1)), int(m.group(2)), int(m.group(3))
        k = sum(a * b for a, b in eval(m.group(4)))
        fl = 2 * B * h * w * cout * k
        extra = f'  {fl / (t / c * 1e-3) / 1e12:7.1f} TFLOP/s/launch'
    print(f'{t:8.3f} ms ({100 * t / tot:5.1f}%) x{c:2d}  {n}{extra}')
cat = collections.OrderedDict()
for n, (c, t) in agg.items():
    if n.startswith('conv'):
        m = re.search(r'(\d+)x(\d+) srcs=(.*)', n)
        taps = [b for a, b in eval(m.group(3))]
        key = f'conv {m.group(1)}px ' + ('3x3' if taps == [9] or taps == [9, 9] else ('3x3+1x1' if 9 in taps else '1x1/other'))
    else:
        key = n.split(':')[0]
    cat.setdefault(key, [0, 0.0]); cat[key][0] += c; cat[key][1] += t
print('--- categories ---')
for k, (c, t) in sorted(cat.items(), key=lambda kv: -kv[1][1]):
    print(f'{t:8.3f} ms ({100 * t / tot:5.1f}%) x{c:3d}  {k}')
