"""torch.profiler breakdown of one train step (kernel time and launch counts by name, CPU ops by count)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
import bench
from ddgan_b200.modules import NCSNpp, Discriminator_small
from ddgan_b200.train import Trainer
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
cfg = bench.train_args()
netG = NCSNpp(cfg).to(dev); netD = Discriminator_small(nc=6, ngf=64, t_emb_dim=256).to(dev)
tr = Trainer(cfg, netG, netD, dev)
real = torch.rand(64, 3, 32, 32, device=dev) * 2 - 1
for i in range(3): tr.step(real, i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(1, 6): tr.step(real, i)
torch.cuda.synchronize()
print('plain step ms (eager)', (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(1, 3): tr.step(real, i)
    torch.cuda.synchronize()
ev = prof.key_averages()
kern = [e for e in ev if e.self_device_time_total > 0 and e.device_type.name == 'CUDA']
tot = sum(e.self_device_time_total for e in kern)
print('total kernel time per step (ms):', tot / 2 / 1e3, ' launches per step:', sum(e.count for e in kern) // 2)
print('--- kernels by time ---')
for e in sorted(kern, key=lambda e: -e.self_device_time_total)[:40]:
    print(f'{e.self_device_time_total/2/1e3:9.3f} ms x{e.count//2:5d}  {e.key[:110]}')
print('--- kernels by launch count ---')
for e in sorted(kern, key=lambda e: -e.count)[:70]:
    print(f'x{e.count//2:5d} {e.self_device_time_total/2/1e3:8.3f} ms  {e.key[:120]}')
print('--- CPU-side ops by count (aten:: / autograd Functions) ---')
cpu = [e for e in ev if e.device_type.name == 'CPU']
for e in sorted(cpu, key=lambda e: -e.count)[:60]:
    print(f'x{e.count//2:5d}  {e.key[:100]}')
