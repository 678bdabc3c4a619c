#!/usr/bin/env python
"""Benchmark of the DDGAN hot path on B200 (contract: prompt section "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU cores (oracle port)

Headline line = workload `cifar10_sample_T4_b64` (BASELINE.json configs[0]): CIFAR-10 NCSN++ (ch 128, ch_mult 1-2-2-2,
nz 100), T = 4 posterior-sampling steps, 64 images per GPU per step, re-randomised weights of the reference architecture,
synthetic noise.  One "step" = one full sampling pass (4 generator forwards + 4 posterior updates) over 64 images per GPU.

The same JSON line carries, as sub-objects with the same fields:
  train          configs[1]: adversarial step, G + Discriminator_small fwd/bwd, lazy R1 (0.02, every 15), Adam, EMA, batch 64/GPU
                 (default schedule skips the generator backward the reference discards; `faithful` = with it)
  bf16           the sampling workload in BF16 mode (precision 1)
  hq256_sample   configs[2]: CelebA-HQ 256 NCSN++ (ch 64, 1-1-2-2-4-4), T = 2, 8 images per GPU
  lsun256_train  configs[3]: LSUN-256 adversarial step with Discriminator_large, R1 gamma 1.0 every 10, batch 8/GPU
  microbench     configs[4] (N = 1 only): upfirdn2d / fused_leaky_relu / GroupNorm sweep vs the measured HBM copy peak
  reference_gpu  (N = 1 only) the reference algorithm in stock eager PyTorch on the same GPU (cuDNN fp32 / TF32 / channels_last)
  cpu_baseline   (N = 1 only) the reference algorithm on the host cores (oracle port), bounded sample of the same workload
`--workload X` runs one of them alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line.  Libraries print there too (NCCL's version banner comes out on file descriptor 1 of every rank),
# so the real stdout is kept on a private descriptor for the result line and descriptor 1 is pointed at stderr for everything else.
sys.stdout.flush()
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit_result(line):
    os.write(_RESULT_FD, (json.dumps(line) + '\n').encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'denoising-diffusion-gan_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

G_FLOP_PER_SAMPLE_FWD = 14.07e9  # SURVEY.md section 8(d), hooked from the reference (conv + linear + NIN + attention)
WORKLOADS = ['all', 'sample', 'train', 'bf16', 'hq256', 'lsun256', 'micro', 'refgpu']


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='all', choices=WORKLOADS + ['both'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step (CIFAR workloads)')
    ap.add_argument('--precision', type=int, default=3, help='3 = BF16x3 (fp32 parity mode, headline), 1 = BF16')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--faithful-wasted-backward', action='store_true',
                    help='train workload only: run the generator backward of the D step whose gradients the reference discards')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget-s', type=float, default=240.0, help='wall-clock cap of the --impl reference run')
    a = ap.parse_args()
    if a.workload == 'both':
        a.workload = 'all'
    return a


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the conv launches of one generator forward, from the committed
    `ncu --set full` capture (profiles/r2_conv_traffic.json, written by tools/ncu_traffic.py); None if absent."""
    p = os.path.join(ROOT, 'profiles', 'r2_conv_traffic.json')
    if os.path.exists(p):
        return json.load(open(p))
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# workload descriptions shared by both arms (the driver compares the `config` objects of the two arms)
# ---------------------------------------------------------------------------------------------------------------------
def sample_config(B):
    return {'workload': 'cifar10_sample_T4_b64', 'model': 'NCSN++ ch128 1-2-2-2 nz100', 'T': 4, 'batch_per_gpu': B,
            'images_per_step_per_gpu': B}


def train_config(B):
    return {'workload': 'cifar10_train_step_b64', 'model': 'NCSN++ ch128 1-2-2-2 nz100 + Discriminator_small ngf64', 'T': 4,
            'batch_per_gpu': B, 'r1_gamma': 0.02, 'lazy_reg': 15, 'optimizer': 'Adam lr_g 1.6e-4 lr_d 1.25e-4 betas (0.5, 0.9)'}


def train_args(hq=False):
    from ddgan_b200 import arch
    if hq:
        # readme.md:41-55 LSUN church 256: ch 64, 1-1-2-2-4-4, T = 4, Discriminator_large, r1_gamma 1.0, lazy_reg 10
        cfg = arch.make_config(image_size=256, num_channels_dae=64, ch_mult=(1, 1, 2, 2, 4, 4), n_mlp=3, num_timesteps=4, ngf=64)
        over = dict(lr_g=1.6e-4, lr_d=1.0e-4, r1_gamma=1.0, lazy_reg=10, ema_decay=0.999)
    else:
        # readme.md:31-37 CIFAR-10 command
        cfg = arch.make_config()
        over = dict(lr_g=1.6e-4, lr_d=1.25e-4, r1_gamma=0.02, lazy_reg=15, ema_decay=0.9999)
    for k, v in dict(beta1_g=0.5, beta2_g=0.9, beta1_d=0.5, beta2_d=0.9, weight_decay_G=0.0, weight_decay_D=0.0,
                     grad_clip_norm=1.0, use_ema=True, **over).items():
        setattr(cfg, k, v)
    return cfg


def randomized_generator_state(cfg):
    """Re-randomised weights of the reference architecture (shapes = NCSNpp(args).state_dict())."""
    import torch
    from ddgan_b200 import arch
    sd = {}
    g = torch.Generator(device='cpu').manual_seed(1)
    for k, shp in arch.ncsnpp_param_shapes(arch.normalize_config(cfg)).items():
        fan = 1
        for d_ in shp[1:]:
            fan *= d_
        if len(shp) >= 2:
            sd[k] = torch.randn(shp, generator=g) * (1.0 / max(fan, 1)) ** 0.5
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.1 + (1.0 if (k.endswith('.weight')) else 0.0)
        if k.endswith('style.bias'):
            sd[k][: shp[0] // 2] += 1.0
    return sd


# ---------------------------------------------------------------------------------------------------------------------
# reference algorithm on the host CPU (oracle port of the reference's CPU path: upfirdn2d_native + torch CPU conv)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sampling(batch, steps, warmup, budget_s=1e9, threads=None):
    import torch
    from oracle import ddgan_oracle as O
    cfg = O.cifar10_config()
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=1)
    pc = O.posterior_coefficients(cfg)
    gen = lambda x, t, z: O.ncsnpp_forward(sd, cfg, x, t, z)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        x = torch.randn(batch, 3, 32, 32)
        t0 = time.perf_counter()
        O.sample_from_model(pc, gen, cfg.num_timesteps, x, cfg.nz)
        times.append(time.perf_counter() - t0)
        # wall-clock cap: keep at least one timed step, then stop when the next one would overrun
        if i >= warmup and (time.perf_counter() - t_start) + times[-1] > budget_s:
            break
    done_warm = min(warmup, max(len(times) - 1, 0))
    timed = times[done_warm:]
    dt = sum(timed) / len(timed)
    return batch / dt, dt, torch.get_num_threads(), len(timed), done_warm


def cpu_train(batch, r1_steps=1, plain_steps=1, warmup=1, threads=None):
    """Reference train-step body (ddgan.py:443-518: D real [+ R1] + fake incl. the generator backward the reference runs there,
    G step; losses and all gradients) through the oracle port; optimiser updates excluded (negligible next to the backward
    passes).  Returns seconds per step with and without the R1 term."""
    import torch
    from oracle import ddgan_oracle as O
    cfg = O.cifar10_config()
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd_g = {k: v.requires_grad_(True) for k, v in O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=1).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in O.randomize_params(O.discriminator_param_shapes(6, 64, 256), seed=2).items()}
    res = {True: [], False: []}
    plan = [False] * warmup + [True] * r1_steps + [False] * plain_steps
    for i, do_r1 in enumerate(plan):
        real = torch.rand(batch, 3, 32, 32) * 2 - 1
        t = torch.randint(0, 4, (batch,))
        noises = [torch.randn(batch, 3, 32, 32) for _ in range(3)]
        z = torch.randn(batch, cfg.nz)
        t0 = time.perf_counter()
        er, gp, ef = O.d_step_losses(sd_g, sd_d, cfg, real, t, noises, z, 0.02, do_r1=do_r1)
        (er + ef + (gp if gp is not None else 0.0)).backward()
        for p_ in list(sd_g.values()) + list(sd_d.values()):
            p_.grad = None
        eg = O.g_step_loss(sd_g, {k: v.detach() for k, v in sd_d.items()}, cfg, real, t, noises, z)
        eg.backward()
        for p_ in sd_g.values():
            p_.grad = None
        if i >= warmup:
            res[do_r1].append(time.perf_counter() - t0)
    s_r1 = sum(res[True]) / max(len(res[True]), 1)
    s_plain = sum(res[False]) / max(len(res[False]), 1)
    return s_r1, s_plain, torch.get_num_threads()


def cpu_train_entry(batch=16):
    s_r1, s_plain, th = cpu_train(batch)
    mix = (s_r1 + 14 * s_plain) / 15.0           # lazy_reg = 15: one R1 step in fifteen
    return {'value': batch / mix, 'unit': 'samples/s', 'cores': th, 'kind': 'port',
            'r1_step_samples_per_s': batch / s_r1, 'plain_step_samples_per_s': batch / s_plain,
            'sample': f'1 R1 step + 1 plain step of the reference train-step body at batch {batch} (after 1 warm-up), {th} threads; '
                      f'value = lazy_reg-15 mix; the reference schedule incl. the discarded generator backward'}


def run_reference(args):
    """Reference arm: the reference's CPU path (oracle port; the reference has no setup.py, so it cannot be pip-installed into
    baseline/_ref -- DESIGN.md section 2) on the host cores, same workload / metric / config as the GPU arm: every step samples
    the full 64 images.  Honours --steps / --warmup up to --cpu-budget-s of wall clock."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    B = args.batch
    ips, dt, th, steps, warm = cpu_sampling(B, args.steps, args.warmup, budget_s=args.cpu_budget_s)
    line = {
        'impl': 'reference', 'metric': 'cifar10_T4_sampled_images_per_sec', 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': sample_config(B),
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': th, 'kind': 'port',
                         'sample': f'{steps} x T=4 sampling of {B} images ({warm} warm-up), {th} threads'},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'reference algorithm (oracle port of the reference CPU path: upfirdn2d_native + torch CPU conv) on the host cores'
                + ('' if steps == args.steps else f'; stopped after {steps} of {args.steps} steps at the {args.cpu_budget_s:.0f} s wall cap'),
    }
    if args.workload in ('all', 'train'):
        e = cpu_train_entry(16)
        line['train'] = {'metric': 'cifar10_train_samples_per_sec', 'value': e['value'], 'unit': 'samples/s',
                         'config': train_config(16), 'cpu_baseline': e}
    emit_result(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group('nccl', device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def timed(self, fn, steps):
        """`steps` calls of fn between barriers, CUDA events on the launching stream; returns ms (this rank)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1)


def conv_roofline(eng, pk, pk_kind, reps=3):
    """Per-launch CUDA events around every tensor-core conv launch of one eager generator forward."""
    import torch
    torch.cuda.synchronize()
    conv_ms = 0.0
    total_ms_eager = 0.0
    for r in range(reps):
        pairs = []
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i, st in enumerate(eng.steps):
            if eng.step_names[i].startswith(('conv', 'attn_fused')):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); st(); b.record()
                pairs.append((a, b))
            else:
                st()
        s1.record()
        torch.cuda.synchronize()
        if r > 0:
            conv_ms += sum(a.elapsed_time(b) for a, b in pairs)
            total_ms_eager += s0.elapsed_time(s1)
    conv_ms /= (reps - 1)
    total_ms_eager /= (reps - 1)
    n_conv = len([n for n in eng.step_names if n.startswith(('conv', 'attn_fused'))])
    achieved = eng.conv_flops / (conv_ms * 1e-3) / 1e12
    peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
    prec3 = eng.prec == 3
    roof = {'bound': 'tensor', 'kernel': 'conv_tc_kernel (all instantiations) + attn_kernel, %d tensor-core launches per generator forward' % n_conv,
            'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
            'peak_source': f'{pk_kind} bf16_tflops_sustained (kernel timed inside a long step)',
            'flops_per_forward': eng.conv_flops, 'conv_ms_per_forward': conv_ms, 'eager_forward_ms': total_ms_eager,
            'launches_per_forward': n_conv,
            'note': ('algorithmic FLOPs (incl. channel padding of the 3-channel input conv); BF16x3 issues 3 MMAs per algorithmic '
                     'MAC, so the fp32-parity mode tops out at 1/3 of the bf16 peak by construction (mma_level_frac = 3 x frac)')
                    if prec3 else 'algorithmic FLOPs; single-pass BF16 operands, fp32 accumulate'}
    if prec3:
        roof['mma_level_frac'] = 3 * achieved / peak
    return roof


def run_sampling(ctx, args, precision, hq=False):
    torch = ctx.torch
    from ddgan_b200 import arch, diffusion
    from ddgan_b200.engine import GeneratorEngine
    if hq:
        cfg = arch.make_config(image_size=256, num_channels_dae=64, ch_mult=(1, 1, 2, 2, 4, 4), n_mlp=3, num_timesteps=2, ngf=64)
        B, S = 8, 256
    else:
        cfg = arch.make_config()
        B, S = args.batch, 32
    T = cfg.num_timesteps
    torch.manual_seed(1024 + ctx.rank)  # seed + rank, ddgan.py:189
    eng = GeneratorEngine(cfg, B, ctx.dev, precision=precision)
    eng.load_state_dict(randomized_generator_state(cfg))
    smp = diffusion.GraphSampler(eng, cfg)
    if not args.no_graph:
        smp.capture()
    x_init = torch.randn(B, 3, S, S, device=ctx.dev)
    launches_per_step = T * (eng.n_launches + 4) + 3
    steps = args.steps if not hq else max(5, min(args.steps, 10))
    for _ in range(max(args.warmup, 3)):
        smp.sample(x_init)
    clocks = ClockSampler(ctx.local).start() if ctx.rank == 0 else None
    ms = ctx.timed(lambda i: smp.sample(x_init), steps)
    clk = clocks.stop() if clocks else None
    # ---- end to end through the public API with host buffers ----
    h_in = torch.randn(B, 3, S, S).pin_memory()
    h_out = torch.empty(B, 3, S, S).pin_memory()
    for _ in range(2):
        h_out.copy_(smp.sample(h_in.to(ctx.dev, non_blocking=True)), non_blocking=True)

    def e2e_step(i):
        h_out.copy_(smp.sample(h_in.to(ctx.dev, non_blocking=True)), non_blocking=True)
    ms_e2e = ctx.timed(e2e_step, steps)
    ms, ms_e2e = ctx.max_over_ranks([ms, ms_e2e])
    line = None
    if ctx.rank == 0:
        pk, pk_kind = peaks()
        images = B * ctx.world * steps
        value = images / (ms * 1e-3)
        roof = conv_roofline(eng, pk, pk_kind)
        tr = conv_traffic()
        if tr is not None and not hq and precision == 3:
            roof['traffic'] = tr['dram_bytes_per_launch']
            roof['algorithmic_bytes_per_launch'] = tr['algorithmic_bytes_per_launch']
            roof['traffic_source'] = tr['source']
        name = 'celebahq256_sample_T2_b8' if hq else 'cifar10_sample_T4_b64'
        conf = ({'workload': name, 'model': 'NCSN++ ch64 1-1-2-2-4-4 nz100 n_mlp3', 'T': T, 'batch_per_gpu': B,
                 'images_per_step_per_gpu': B} if hq else sample_config(B))
        flop_fwd = eng.conv_flops / B
        line = {
            'metric': ('celebahq256_T2' if hq else 'cifar10_T4') + '_sampled_images_per_sec', 'value': value, 'unit': 'images/s',
            'n_gpus': ctx.world, 'steps': steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (BF16x3 split operands on tcgen05, fp32 accumulate)' if precision == 3 else 'bf16',
            'data': 'synthetic', 'config': conf,
            'run': {'parallelism': f'batch-partitioned x{ctx.world}, no collective', 'cuda_graph': not args.no_graph,
                    'l2': 'working set per step (GBs of activations + packed weights) exceeds the 126 MB L2; no explicit flush'},
            'e2e': {'value': images / (ms_e2e * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': B * 3 * S * S * 4,
                    'd2h_bytes_per_step': B * 3 * S * S * 4},
            'gpu_launches': launches_per_step * steps, 'clocks': clk, 'roofline': roof,
            'model_tflops': T * flop_fwd * value / 1e12,
        }
    del smp, eng
    torch.cuda.empty_cache()
    return line


def run_train(ctx, args, hq=False, faithful=False, profile=True):
    torch, dist = ctx.torch, ctx.dist
    from ddgan_b200 import ops
    from ddgan_b200.modules import NCSNpp, Discriminator_small, Discriminator_large
    from ddgan_b200.train import Trainer
    cfg = train_args(hq)
    B = 8 if hq else args.batch
    S = cfg.image_size
    torch.manual_seed(1024 + ctx.rank)
    netG = NCSNpp(cfg).to(ctx.dev)
    D = Discriminator_large if hq else Discriminator_small
    netD = D(nc=2 * cfg.num_channels, ngf=cfg.ngf, t_emb_dim=cfg.t_emb_dim).to(ctx.dev)
    netG.precision = netD.precision = args.precision
    tr = Trainer(cfg, netG, netD, ctx.dev, distributed=ctx.world > 1, skip_discarded_g_backward=not faithful)
    real = (torch.rand(B, 3, S, S, device=ctx.dev) * 2 - 1)
    use_graph = not args.no_graph
    steps = max(cfg.lazy_reg, min(args.steps, 2 * cfg.lazy_reg)) if not hq else cfg.lazy_reg
    if use_graph:
        tr.capture((B, 3, S, S), warmup=3)
        step_fn = tr.step_graphed
    else:
        step_fn = tr.step
        for i in range(3):
            tr.step(real, 0 if i == 0 else i)      # the first warm-up step exercises the R1 double-backward path
    for i in range(2):
        step_fn(real, i)
    clocks = ClockSampler(ctx.local).start() if ctx.rank == 0 else None
    ms = ctx.timed(lambda i: step_fn(real, i), steps)      # lazy R1 on steps 0, lazy_reg, ... as in the reference
    clk = clocks.stop() if clocks else None
    # end to end: host batch -> device every step, both losses read back (.item() as ddgan.py:480,510)
    h_real = (torch.rand(B, 3, S, S) * 2 - 1).pin_memory()

    def e2e_step(i):
        errD, errG = step_fn(h_real.to(ctx.dev, non_blocking=True), i)
        _ = errD.item(), errG.item()
    ms_e2e = ctx.timed(e2e_step, steps)
    ms, ms_e2e = ctx.max_over_ranks([ms, ms_e2e])
    # tensor-core kernel time inside one R1 step and one plain step (per-launch CUDA events on the launching stream).
    # Every rank runs these two steps (they contain the gradient all-reduce); only rank 0 reports.
    agg = {}
    if profile:
        ops.PROFILE['on'] = True
        for gs in (cfg.lazy_reg, cfg.lazy_reg + 1):
            ops.PROFILE['records'] = []
            tr.step(real, gs)
            torch.cuda.synchronize()
            for name, fl, a, b in ops.PROFILE['records']:
                e = agg.setdefault((gs, name), [0, 0.0, 0.0])
                e[0] += 1; e[1] += fl; e[2] += a.elapsed_time(b)
        ops.PROFILE['on'] = False
    ctx.barrier()
    line = None
    if ctx.rank == 0:
        L = cfg.lazy_reg
        samples = B * ctx.world * steps
        value = samples / (ms * 1e-3)
        line = {
            'metric': ('lsun256' if hq else 'cifar10') + '_train_samples_per_sec', 'value': value, 'unit': 'samples/s',
            'n_gpus': ctx.world, 'steps': steps, 'warmup': 5, 'ms_per_step': ms / steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 (BF16x3 split operands on tcgen05, fp32 accumulate)' if args.precision == 3 else 'bf16',
            'data': 'synthetic',
            'config': ({'workload': 'lsun256_train_step_b8', 'model': 'NCSN++ ch64 1-1-2-2-4-4 nz100 + Discriminator_large ngf64', 'T': 4,
                        'batch_per_gpu': B, 'r1_gamma': 1.0, 'lazy_reg': 10, 'optimizer': 'Adam lr_g 1.6e-4 lr_d 1e-4 betas (0.5, 0.9)'}
                       if hq else train_config(B)),
            'run': {'parallelism': f'data parallel x{ctx.world}: one flat NCCL all-reduce(sum) per network per step, 1/world in the optimiser pass',
                    'cuda_graph': use_graph,
                    'd_step_generator_backward': 'computed (reference-faithful)' if faithful else
                    'skipped: those G gradients are zeroed by netG.zero_grad() (ddgan.py:489) before any use; parameter updates identical',
                    'l2': 'activations saved for backward (GBs per step) exceed the 126 MB L2; no explicit flush'},
            'e2e': {'value': samples / (ms_e2e * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': B * 3 * S * S * 4, 'd2h_bytes_per_step': 8},
            'clocks': clk,
        }
        if profile:
            def mix(name, idx):
                return (agg.get((L, name), [0, 0, 0])[idx] + (L - 1) * agg.get((L + 1, name), [0, 0, 0])[idx]) / float(L)
            conv_ms, conv_fl = mix('conv_tc', 2) + mix('wgrad_tc', 2), mix('conv_tc', 1) + mix('wgrad_tc', 1)
            n_launch = mix('conv_tc', 0) + mix('wgrad_tc', 0)
            pk, pk_kind = peaks()
            peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
            achieved = conv_fl / (conv_ms * 1e-3) / 1e12
            line['gpu_launches'] = int(n_launch * steps)
            line['roofline'] = {'bound': 'tensor', 'kernel': f'conv_tc_kernel + wgrad_tc_kernel (fwd, dgrad, wgrad; lazy-R1 mix 1:{L - 1})',
                                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                                'peak_source': f'{pk_kind} bf16_tflops_sustained', 'tc_ms_per_step': conv_ms, 'tc_flops_per_step': conv_fl,
                                'tc_launches_per_step': n_launch, 'tc_share_of_step': conv_ms / (ms / steps),
                                'note': 'algorithmic FLOPs of the launched GEMMs (padded channels included); BF16x3 issues 3 MMAs per MAC'}
            # launched tensor-core FLOPs per second of wall clock (not the reference schedule's 108 GFLOP/sample)
            line['model_tflops'] = conv_fl / (ms / steps * 1e-3) / 1e12
    del tr, netG, netD
    torch.cuda.empty_cache()
    return line


def run_microbench(ctx):
    """BASELINE configs[4], bounded: upfirdn2d down/up x2, fused_leaky_relu, AdaGN+SiLU forward and GroupNorm backward at the
    CIFAR sizes the path runs plus one large size each, vs the measured HBM copy peak; L2 flushed by reading a 512 MB buffer
    between timed launches (tools/microbench.py is the full sweep)."""
    torch = ctx.torch
    from ddgan_b200 import ops
    pk, pk_kind = peaks()
    peak = pk['hbm_gbs']
    dev = ctx.dev
    flush = torch.zeros(512 * 1024 * 1024 // 4, device=dev)

    def timeit(fn, reps=4):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps
    k4 = torch.tensor([1., 3., 3., 1.]); k4 = torch.outer(k4, k4); k4 = (k4 / k4.sum()).to(dev)
    rows = []
    N = 64
    for C, H in ((128, 32), (256, 32), (256, 64), (128, 128)):
        numel = N * C * H * H
        x = torch.randn(N * C, H, H, device=dev)
        x4 = x.view(N, C, H, H)
        rows.append(('upfirdn2d_down2', C, H, 4 * numel * 1.25, timeit(lambda: ops.upfirdn2d_raw(x, k4, 1, 1, 2, 2, 1, 1, 1, 1))))
        rows.append(('upfirdn2d_up2', C, H, 4 * numel * 5.0, timeit(lambda: ops.upfirdn2d_raw(x, k4 * 4, 2, 2, 1, 1, 2, 1, 2, 1))))
        b = torch.randn(C, device=dev)
        rows.append(('fused_leaky_relu', C, H, 8.0 * numel, timeit(lambda: ops.fused_bias_act(x4, b, None, 3, 0, 0.2, 2 ** 0.5))))
        xb = x.to(torch.bfloat16); xb4 = xb.view(N, C, H, H)
        # 16-bit I/O (fp16 / bf16 tensors are accepted as in the reference): register-tiled x2 kernels, fp32 taps and accumulation
        rows.append(('upfirdn2d_down2_bf16', C, H, 2 * numel * 1.25, timeit(lambda: ops.upfirdn2d_raw(xb, k4, 1, 1, 2, 2, 1, 1, 1, 1))))
        rows.append(('upfirdn2d_up2_bf16', C, H, 2 * numel * 5.0, timeit(lambda: ops.upfirdn2d_raw(xb, k4 * 4, 2, 2, 1, 1, 2, 1, 2, 1))))
        rows.append(('fused_leaky_relu_bf16', C, H, 4.0 * numel, timeit(lambda: ops.fused_bias_act(xb4, b, None, 3, 0, 0.2, 2 ** 0.5))))
        del xb, xb4
        G = min(C // 4, 32)
        gamma = torch.randn(N, C, device=dev); beta = torch.randn(N, C, device=dev)
        rows.append(('adagn_silu_fwd', C, H, 8.0 * numel, timeit(lambda: ops.groupnorm_fwd(x4, G, gamma, beta, per_sample=True, act=ops.ACT_SILU))))
        y, mean, rstd = ops.groupnorm_fwd(x4, G, gamma, beta, per_sample=True, act=ops.ACT_SILU)
        dy = torch.randn_like(x4)
        rows.append(('adagn_silu_bwd', C, H, 12.0 * numel,
                     timeit(lambda: ops.groupnorm_bwd(x4, dy, G, mean, rstd, gamma, beta, per_sample=True, act=ops.ACT_SILU))))
        del x, x4, y, dy
    out = [{'op': r[0], 'C': r[1], 'HW': r[2], 'N': N, 'dtype': 'bf16' if r[0].endswith('_bf16') else 'f32', 'us': r[4] * 1e3, 'GBps': r[3] / r[4] / 1e6,
            'frac_of_hbm_peak': r[3] / r[4] / 1e6 / peak} for r in rows]
    return {'peak_GBps': peak, 'peak_source': f'{pk_kind} hbm_gbs (copy)', 'l2': 'flushed by a 512 MB read between launches',
            'bytes_model': 'upfirdn2d e(in+out); fused_leaky_relu 2e/elem; AdaGN+SiLU fwd 8/elem; bwd 12/elem (x, dy in; dx out); e = 4 (f32) / 2 (bf16)',
            'rows': out}


def run_reference_gpu(ctx, args):
    """The reference algorithm (the oracle's functional restatement of NCSNpp.forward + sample_posterior) in stock eager PyTorch
    on this GPU -- i.e. the reference's own GPU path: cuDNN convolutions and ATen elementwise kernels -- for the sampling
    workload: cuDNN fp32, cuDNN TF32 (PyTorch's default for convolutions) and TF32 with channels_last weights/activations."""
    torch = ctx.torch
    from oracle import ddgan_oracle as O
    cfg = O.cifar10_config()
    B = args.batch
    dev = ctx.dev
    sd = {k: v.to(dev) for k, v in O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=1).items()}
    pc = O.posterior_coefficients(cfg)
    for k, v in list(vars(pc).items()):
        if torch.is_tensor(v):
            setattr(pc, k, v.to(dev))
    res = {}
    import torch.nn.functional as F

    # The oracle states GroupNorm (float64 statistics) and upfirdn2d (16 shifted adds) explicitly for checking; the reference's
    # GPU path runs nn.GroupNorm and a conv-shaped FIR, so this leg swaps in the stock ATen / cuDNN formulations of the two.
    def gn_aten(x, groups, eps=1e-6, weight=None, bias=None):
        return F.group_norm(x, groups, weight, bias, eps)

    def upfirdn2d_conv(x, kernel, up=1, down=1, pad=(0, 0)):
        n, c, h, w = x.shape
        kh, kw = kernel.shape
        u = x.reshape(n * c, 1, h, w)
        if up > 1:
            z = u.new_zeros(n * c, 1, h * up, w * up)
            z[:, :, ::up, ::up] = u
            u = z
        p0, p1 = pad
        u = F.pad(u, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
        u = u[:, :, max(-p0, 0): u.shape[2] - max(-p1, 0), max(-p0, 0): u.shape[3] - max(-p1, 0)]
        y = F.conv2d(u, torch.flip(kernel, [0, 1]).view(1, 1, kh, kw))
        return y[:, :, ::down, ::down].reshape(n, c, (y.shape[2] + down - 1) // down, (y.shape[3] + down - 1) // down)
    saved = (O.group_norm, O.upfirdn2d)
    O.group_norm, O.upfirdn2d = gn_aten, upfirdn2d_conv

    def sample(sd_, cl):
        x = torch.randn(B, 3, 32, 32, device=dev)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            for i in reversed(range(cfg.num_timesteps)):
                t = torch.full((B,), i, dtype=torch.int64, device=dev)
                z = torch.randn(B, cfg.nz, device=dev)
                x0 = O.ncsnpp_forward(sd_, cfg, x, t, z)
                x = O.sample_posterior(pc, x0, x, t, torch.randn_like(x))
        return x
    old = torch.backends.cudnn.allow_tf32
    try:
        for name, tf32, cl in (('cudnn_fp32', False, False), ('cudnn_tf32', True, False), ('cudnn_tf32_channels_last', True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            sd_ = {k: (v.contiguous(memory_format=torch.channels_last) if (cl and v.dim() == 4) else v) for k, v in sd.items()}
            for _ in range(2):
                sample(sd_, cl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                sample(sd_, cl)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res[name] = {'value': B / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms}
    finally:
        torch.backends.cudnn.allow_tf32 = old
        O.group_norm, O.upfirdn2d = saved
    res['note'] = ('eager PyTorch (cuDNN + ATen) running the reference algorithm on the same GPU, 64 images, T = 4; TF32 is what the '
                   'reference gets by default and does not meet the 1e-4 parity gate (SURVEY.md section 7)')
    return res


def run_b200(args):
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    wl = args.workload
    line = None
    single = ctx.world == 1
    if wl in ('all', 'sample'):
        line = run_sampling(ctx, args, args.precision)
    if wl == 'train':
        line = run_train(ctx, args, faithful=args.faithful_wasted_backward)
    elif wl == 'all':
        tl = run_train(ctx, args)
        fl = run_train(ctx, args, faithful=True, profile=False)
        if ctx.rank == 0:
            tl['faithful'] = {'value': fl['value'], 'unit': 'samples/s', 'ms_per_step': fl['ms_per_step'], 'e2e': fl['e2e'],
                              'note': 'reference-faithful schedule: the D step also back-propagates through the generator '
                                      '(gradients discarded by ddgan.py:489)'}
            line['train'] = tl
    def sub(key, fn):
        """Secondary workloads ride in the headline line; a failure there is recorded, not fatal (it is deterministic across
        ranks: every rank takes the same branch)."""
        try:
            r = fn()
        except Exception as e:  # noqa: BLE001
            r = {'error': f'{type(e).__name__}: {e}'[:300]}
            torch.cuda.synchronize()
        if ctx.rank == 0:
            line[key] = r
    if wl == 'bf16':
        line = run_sampling(ctx, args, 1)
    elif wl == 'all' and args.precision != 1:
        sub('bf16', lambda: run_sampling(ctx, args, 1))
    if wl == 'hq256':
        line = run_sampling(ctx, args, args.precision, hq=True)
    elif wl == 'all':
        sub('hq256_sample', lambda: run_sampling(ctx, args, args.precision, hq=True))
    if wl == 'lsun256':
        line = run_train(ctx, args, hq=True)
    elif wl == 'all':
        sub('lsun256_train', lambda: run_train(ctx, args, hq=True, profile=False))
    # ---- the legs below burn host CPU or a single GPU: only in single-process runs (never behind an NCCL barrier) ----
    if single and wl in ('all', 'micro'):
        m = run_microbench(ctx)
        if wl == 'micro':
            line = {'metric': 'microbench', 'microbench': m}
        else:
            line['microbench'] = m
    if single and wl in ('all', 'refgpu'):
        r = run_reference_gpu(ctx, args)
        if wl == 'refgpu':
            line = {'metric': 'reference_gpu', 'reference_gpu': r}
        else:
            line['reference_gpu'] = r
    if single and not args.skip_cpu_baseline and wl in ('all', 'sample', 'train'):
        torch.cuda.synchronize()
        if wl in ('all', 'sample'):
            ips, dt, th, st, wm = cpu_sampling(args.batch, 1, 1)
            line['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': th, 'kind': 'port',
                                    'sample': f'1 x T=4 sampling of {args.batch} images (after 1 warm-up pass), {th} threads'}
        if wl in ('all', 'train'):
            e = cpu_train_entry(16)
            (line['train'] if wl == 'all' else line)['cpu_baseline'] = e
    if ctx.rank == 0 and line is not None:
        emit_result(line)
    if ctx.world > 1:
        dbg = os.environ.get('DDG_BENCH_DEBUG')
        if dbg:
            print(f'[rank {ctx.rank}] before final barrier', file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        dist.barrier()
        if dbg:
            print(f'[rank {ctx.rank}] after final barrier', file=sys.stderr, flush=True)
        # The result line is out; tearing NCCL down must not be able to hang the run (seen with collectives captured in CUDA graphs
        # on a side stream while the graphs were still alive): a watchdog ends the process after 20 s.
        import gc
        gc.collect()
        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        if dbg:
            print(f'[rank {ctx.rank}] destroyed', file=sys.stderr, flush=True)
        os._exit(0)


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
