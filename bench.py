#!/usr/bin/env python
"""Benchmark of the DDGAN hot path on B200 (contract: prompt section "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU cores (oracle port)

Workload `cifar10_sample_T4_b64`: BASELINE.json configs[0] -- CIFAR-10 NCSN++ (ch 128, ch_mult 1-2-2-2, nz 100), T = 4
posterior-sampling steps, 64 images per GPU per step, random-init-shaped (re-randomised) weights, synthetic noise.
One "step" = one full sampling pass (4 generator forwards + 4 posterior updates) over one batch of 64 images per GPU.
The default run (`--workload both`) also times the adversarial train step (configs[1]: G + Discriminator_small fwd/bwd, lazy R1,
Adam, EMA at batch 64/GPU) and reports it under the key `train` of the same JSON line; `--workload train` prints it alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'denoising-diffusion-gan_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

G_FLOP_PER_SAMPLE_FWD = 14.07e9  # SURVEY.md section 8(d), hooked from the reference (conv + linear + NIN + attention)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='both', choices=['both', 'sample', 'train'],
                    help="'both' (default): sampling line with the train-step results under key 'train'")
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--precision', type=int, default=3, help='3 = BF16x3 (fp32 parity mode, headline), 1 = BF16')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--faithful-wasted-backward', action='store_true',
                    help='also run the generator backward of the D step whose gradients the reference discards (ddgan.py:489)')
    ap.add_argument('--cpu-sample-batch', type=int, default=8)
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


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


def cpu_baseline_sampling(batch, steps, warmup, threads=None):
    """Reference algorithm on the host CPU: the oracle port (upfirdn2d as in upfirdn2d_native + torch CPU conv), all cores."""
    import torch
    from oracle import ddgan_oracle as O
    cfg = O.cifar10_config()
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=1)
    pc = O.posterior_coefficients(cfg)
    gen = lambda x, t, z: O.ncsnpp_forward(sd, cfg, x, t, z)
    times = []
    for i in range(warmup + steps):
        x = torch.randn(batch, 3, 32, 32)
        t0 = time.perf_counter()
        O.sample_from_model(pc, gen, cfg.num_timesteps, x, cfg.nz)
        times.append(time.perf_counter() - t0)
    times = times[warmup:]
    dt = sum(times) / len(times)
    return batch / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    ips, dt, th = cpu_baseline_sampling(args.cpu_sample_batch, steps, warm)
    line = {
        'impl': 'reference', 'metric': 'cifar10_T4_sampled_images_per_sec', 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'cifar10_sample_T4_b64', 'model': 'NCSN++ ch128 1-2-2-2 nz100', 'T': 4,
                   'note': f'reference algorithm (oracle port of the reference CPU path) on host cores; each step samples '
                           f'{args.cpu_sample_batch} images (bounded sample of the 64-image workload)'},
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': th, 'kind': 'port',
                         'sample': f'{steps} x T=4 sampling of {args.cpu_sample_batch} images, {th} threads'},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if args.workload in ('both', 'train'):
        sps, dt2, th2 = cpu_baseline_train(4, 1, 1)
        line['train'] = {'metric': 'cifar10_train_samples_per_sec', 'value': sps, 'unit': 'samples/s', 'ms_per_step': dt2 * 1e3,
                         'cpu_baseline': {'value': sps, 'unit': 'samples/s', 'cores': th2, 'kind': 'port',
                                          'sample': '1 train step (D real + fake, G; no R1) at batch 4 after a warm-up step with R1'}}
    print(json.dumps(line), flush=True)


def cpu_baseline_train(batch, steps, warmup, threads=None):
    """Reference train-step body (ddgan.py:443-518: D real + R1 + fake, G step; losses and all gradients) on the host CPU via
    the oracle port; optimiser updates excluded (negligible next to the backward passes)."""
    import torch
    from oracle import ddgan_oracle as O
    cfg = O.cifar10_config()
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd_g = {k: v.requires_grad_(True) for k, v in O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=1).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in O.randomize_params(O.discriminator_param_shapes(6, 64, 256), seed=2).items()}
    times = []
    for i in range(warmup + steps):
        real = torch.rand(batch, 3, 32, 32) * 2 - 1
        t = torch.randint(0, 4, (batch,))
        noises = [torch.randn(batch, 3, 32, 32) for _ in range(3)]
        z = torch.randn(batch, cfg.nz)
        t0 = time.perf_counter()
        er, gp, ef = O.d_step_losses(sd_g, sd_d, cfg, real, t, noises, z, 0.02, do_r1=(i % 15 == 0))
        (er + ef + (gp if gp is not None else 0.0)).backward()
        for p_ in list(sd_g.values()) + list(sd_d.values()):
            p_.grad = None
        eg = O.g_step_loss(sd_g, {k: v.detach() for k, v in sd_d.items()}, cfg, real, t, noises, z)
        eg.backward()
        for p_ in sd_g.values():
            p_.grad = None
        times.append(time.perf_counter() - t0)
    times = times[warmup:]
    dt = sum(times) / len(times)
    return batch / dt, dt, torch.get_num_threads()


def train_args():
    from ddgan_b200 import arch
    cfg = arch.make_config()
    # readme.md:31-37 CIFAR-10 command; betas = the fork's flag defaults (train_ddgan.py:86-89)
    for k, v in dict(lr_g=1.6e-4, lr_d=1.25e-4, beta1=0.5, beta2=0.9, r1_gamma=0.02, lazy_reg=15, grad_clip_norm=1.0,
                     ema_decay=0.9999, use_ema=True, batch_size=64).items():
        setattr(cfg, k, v)
    return cfg


def run_b200_train(args, emit=True):
    import torch
    import torch.distributed as dist
    from ddgan_b200 import ops
    from ddgan_b200.modules import NCSNpp, Discriminator_small
    from ddgan_b200.train import Trainer
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)
    cfg = train_args()
    B = args.batch
    torch.manual_seed(1024 + rank)
    netG = NCSNpp(cfg).to(dev)
    netD = Discriminator_small(nc=2 * cfg.num_channels, ngf=cfg.ngf, t_emb_dim=cfg.t_emb_dim).to(dev)
    netG.precision = netD.precision = args.precision
    tr = Trainer(cfg, netG, netD, dev, distributed=world > 1, skip_discarded_g_backward=not args.faithful_wasted_backward)
    real = (torch.rand(B, 3, 32, 32, device=dev) * 2 - 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph
    if use_graph:
        tr.capture((B, 3, 32, 32), warmup=max(args.warmup, 3))
        step_fn = tr.step_graphed
    else:
        step_fn = tr.step
        for i in range(max(args.warmup, 3)):
            tr.step(real, 0 if i == 0 else i)      # the first warm-up step exercises the R1 double-backward path
    for i in range(2):
        step_fn(real, i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_fn(real, i)                            # lazy R1 on steps 0, 15, 30, ... as in the reference
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    # end to end: host batch -> device every step, both losses read back (.item() as ddgan.py:480,510)
    h_real = (torch.rand(B, 3, 32, 32) * 2 - 1).pin_memory()
    barrier()
    e0.record()
    for i in range(args.steps):
        errD, errG = step_fn(h_real.to(dev, non_blocking=True), i)
        _ = errD.item(), errG.item()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(tt[0]), float(tt[1])
    # tensor-core kernel time inside one R1 step and one plain step (per-launch CUDA events on the launching stream).
    # Every rank runs these two steps (they contain the gradient all-reduce); only rank 0 reports.
    ops.PROFILE['on'] = True
    agg = {}
    for gs in (15, 16):
        ops.PROFILE['records'] = []
        tr.step(real, gs)
        torch.cuda.synchronize()
        for name, fl, a, b in ops.PROFILE['records']:
            e = agg.setdefault((gs, name), [0, 0.0, 0.0])
            e[0] += 1; e[1] += fl; e[2] += a.elapsed_time(b)
    ops.PROFILE['on'] = False
    barrier()
    line = None
    if rank == 0:
        # weighted by the lazy_reg mix: 1 R1 step + 14 plain steps
        def mix(name, idx):
            return (agg.get((15, name), [0, 0, 0])[idx] + 14 * agg.get((16, name), [0, 0, 0])[idx]) / 15.0
        conv_ms, conv_fl = mix('conv_tc', 2) + mix('wgrad_tc', 2), mix('conv_tc', 1) + mix('wgrad_tc', 1)
        n_launch = mix('conv_tc', 0) + mix('wgrad_tc', 0)
        pk, pk_kind = peaks()
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12
        samples = B * world * args.steps
        value = samples / (ms * 1e-3)
        line = {
            'metric': 'cifar10_train_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 (BF16x3 split operands on tcgen05, fp32 accumulate)' if args.precision == 3 else 'bf16',
            'data': 'synthetic',
            'config': {'workload': 'cifar10_train_step_b64', 'model': 'NCSN++ ch128 1-2-2-2 nz100 + Discriminator_small ngf64',
                       'T': 4, 'batch_per_gpu': B, 'r1_gamma': 0.02, 'lazy_reg': 15, 'optimizer': 'Adam lr_g 1.6e-4 lr_d 1.25e-4',
                       'parallelism': f'data parallel x{world}: one flat NCCL all-reduce(mean) per network per step',
                       'cuda_graph': use_graph,
                       'd_step_generator_backward': 'computed (reference-faithful)' if args.faithful_wasted_backward else
                       'skipped: those G gradients are zeroed by netG.zero_grad() (ddgan.py:489) before any use; parameter updates identical',
                       'l2': 'activations saved for backward (~10 GB per step) exceed the 126 MB L2; no explicit flush'},
            'e2e': {'value': samples / (ms_e2e * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': B * 3 * 32 * 32 * 4, 'd2h_bytes_per_step': 8},
            'gpu_launches': int(n_launch * args.steps),
            'clocks': clk,
            'roofline': {'bound': 'tensor', 'kernel': 'conv_tc_kernel + wgrad_tc_kernel (fwd, dgrad, wgrad; lazy-R1 mix 1:14)',
                         'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                         'peak_source': f'{pk_kind} bf16_tflops_sustained', 'tc_ms_per_step': conv_ms, 'tc_flops_per_step': conv_fl,
                         'tc_launches_per_step': n_launch,
                         'note': 'algorithmic FLOPs of the launched GEMMs (padded channels included); BF16x3 issues 3 MMAs per MAC'},
            'model_tflops': 108e9 * value / 1e12,
        }
        if not args.skip_cpu_baseline:
            sps, dt, th = cpu_baseline_train(4, 1, 1)
            line['cpu_baseline'] = {'value': sps, 'unit': 'samples/s', 'cores': th, 'kind': 'port',
                                    'sample': f'1 train step (D real + fake, G; no R1) at batch 4 after 1 warm-up step with R1, {th} threads'}
        if emit:
            print(json.dumps(line), flush=True)
    if world > 1 and emit:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ddgan_b200 import arch, diffusion
    from ddgan_b200.engine import GeneratorEngine
    from ddgan_b200 import ops

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = arch.make_config()
    B = args.batch
    torch.manual_seed(1024 + rank)  # seed + rank, ddgan.py:189
    # re-randomised weights of the reference architecture (shapes = NCSNpp(args).state_dict())
    sd = {}
    g = torch.Generator(device='cpu').manual_seed(1)
    for k, shp in arch.ncsnpp_param_shapes(cfg).items():
        fan = 1
        for d_ in shp[1:]:
            fan *= d_
        if len(shp) >= 2:
            sd[k] = torch.randn(shp, generator=g) * (1.0 / max(fan, 1)) ** 0.5
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.1 + (1.0 if (k.endswith('.weight')) else 0.0)
        if k.endswith('style.bias'):
            sd[k][: shp[0] // 2] += 1.0
    eng = GeneratorEngine(cfg, B, dev, precision=args.precision)
    eng.load_state_dict(sd)
    smp = diffusion.GraphSampler(eng, cfg)
    if not args.no_graph:
        smp.capture()
    x_init = torch.randn(B, 3, 32, 32, device=dev)
    launches_per_step = cfg.num_timesteps * (eng.n_launches + 4) + 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    line = None
    # ---- device-resident throughput ----
    for _ in range(max(args.warmup, 3)):
        smp.sample(x_init)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        smp.sample(x_init)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    # ---- end to end through the public API with host buffers ----
    h_in = torch.randn(B, 3, 32, 32).pin_memory()
    h_out = torch.empty(B, 3, 32, 32).pin_memory()
    for _ in range(2):
        h_out.copy_(smp.sample(h_in.to(dev, non_blocking=True)), non_blocking=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        xi = h_in.to(dev, non_blocking=True)
        h_out.copy_(smp.sample(xi), non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        # ---- roofline of the dominant kernel (conv_tc_kernel): per-launch CUDA events on the launching stream ----
        conv_idx = [i for i, n in enumerate(eng.step_names) if n.startswith('conv') or '_attn' in n]
        evs = []
        torch.cuda.synchronize()
        reps = 3
        conv_ms = 0.0
        total_ms_eager = 0.0
        for r in range(reps):
            pairs = []
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i, st in enumerate(eng.steps):
                if eng.step_names[i].startswith('conv'):
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
        n_conv = len([n for n in eng.step_names if n.startswith('conv')])
        pk, pk_kind = peaks()
        achieved = eng.conv_flops / (conv_ms * 1e-3) / 1e12
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        images = B * world * args.steps
        value = images / (ms * 1e-3)
        e2e = images / (ms_e2e * 1e-3)
        line = {
            'metric': 'cifar10_T4_sampled_images_per_sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 (BF16x3 split operands on tcgen05, fp32 accumulate)' if args.precision == 3 else 'bf16',
            'data': 'synthetic',
            'config': {'workload': 'cifar10_sample_T4_b64', 'model': 'NCSN++ ch128 1-2-2-2 nz100', 'T': 4, 'batch_per_gpu': B,
                       'parallelism': f'batch-partitioned x{world}, no collective', 'cuda_graph': not args.no_graph,
                       'l2': 'working set per step (~4 GB activations + 0.4 GB packed weights) exceeds the 126 MB L2; no explicit flush'},
            'e2e': {'value': e2e, 'unit': 'images/s', 'h2d_bytes_per_step': B * 3 * 32 * 32 * 4, 'd2h_bytes_per_step': B * 3 * 32 * 32 * 4},
            'gpu_launches': launches_per_step * args.steps,
            'clocks': clk,
            'roofline': {'bound': 'tensor', 'kernel': 'conv_tc_kernel (all instantiations, %d launches per generator forward)' % n_conv,
                         'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                         'peak_source': f'{pk_kind} bf16_tflops_sustained (kernel timed inside a long step)',
                         'flops_per_forward': eng.conv_flops, 'conv_ms_per_forward': conv_ms, 'eager_forward_ms': total_ms_eager,
                         'note': 'algorithmic FLOPs (incl. channel padding of the 3-channel input conv); BF16x3 issues 3 MMAs per '
                                 'algorithmic MAC, so the fp32-parity mode tops out near 1/3 of the bf16 peak by construction'},
            'model_tflops': 4 * G_FLOP_PER_SAMPLE_FWD * value / 1e12,
        }
        if not args.skip_cpu_baseline:
            ips, dt, th = cpu_baseline_sampling(args.cpu_sample_batch, 1, 1)
            line['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': th, 'kind': 'port',
                                    'sample': f'1 x T=4 sampling of {args.cpu_sample_batch} images (after 1 warm-up), {th} threads'}
    if args.workload == 'both':
        del smp, eng
        torch.cuda.empty_cache()
        targs = argparse.Namespace(**vars(args))
        targs.steps = max(15, min(args.steps, 30))
        tl = run_b200_train(targs, emit=False)
        if rank == 0:
            line['train'] = tl
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'train':
        run_b200_train(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
