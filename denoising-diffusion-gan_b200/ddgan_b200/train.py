"""The adversarial training step of DDGAN (ddgan.py:443-518) and its data-parallel glue (ddgan.py:30-33, 292-294, 363-365).

`Trainer.step(real, global_step)` is the loop body of the reference: D step (real, lazy R1 every `lazy_reg`, fake), clip,
Adam; G step, clip, Adam; EMA.  Data parallelism follows the rule SURVEY.md section 2.4 established empirically for the
reference's DDP: every rank accumulates all of a network's gradient contributions locally and ONE all-reduce(mean) per
network per step makes them identical on all ranks -- implemented here as a single flat-buffer NCCL all-reduce per network
(`FlatGradAllReducer`) instead of DDP's per-bucket hooks around three `backward()` calls."""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import diffusion


def broadcast_params(params, src=0):
    """ddgan.py:30-33."""
    for p in params:
        dist.broadcast(p.data, src=src)


class FlatGradAllReducer:
    """All-reduce(mean) of the accumulated .grad of a parameter list through one flat buffer (one collective per call)."""

    def __init__(self, params, world_size=None, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, device=p0.device, dtype=p0.dtype)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def allreduce(self):
        if self.world <= 1:
            return
        with torch.no_grad():
            for v, p in zip(self.views, self.params):
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(self.world)
            for v, p in zip(self.views, self.params):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    p.grad.copy_(v)


class EMA:
    """ema.py:45-55: ema = decay * ema + (1 - decay) * param, per parameter (multi-tensor ops)."""

    def __init__(self, model, ema_decay=0.999):
        self.decay = ema_decay
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.names = [n for n, p in model.named_parameters() if p.requires_grad]
        self.shadow = [p.detach().clone() for p in self.params]

    @torch.no_grad()
    def step(self):
        if self.decay <= 0:
            return
        torch._foreach_mul_(self.shadow, self.decay)
        torch._foreach_add_(self.shadow, [p.detach() for p in self.params], alpha=1.0 - self.decay)

    def state_dict(self):
        return {n: s.cpu() for n, s in zip(self.names, self.shadow)}

    @torch.no_grad()
    def swap_parameters_with_ema(self, store_params_in_ema=True):
        for p, s in zip(self.params, self.shadow):
            if store_params_in_ema:
                tmp = p.detach().clone()
                p.data.copy_(s)
                s.copy_(tmp)
            else:
                p.data.copy_(s)


class FlatAdam:
    """clip_grad_norm_ + Adam (+ EMA) over flat parameter / gradient arenas: 3 launches per step instead of ~40 foreach
    kernels over hundreds of tensors.  The module's parameters and their .grad become views into the arenas, so the DP
    all-reduce runs on the gradient arena directly (no flatten / unflatten copies)."""

    def __init__(self, module, lr, betas=(0.5, 0.9), eps=1e-8, weight_decay=0.0, max_norm=0.0, ema_decay=0.0):
        from . import _lib
        self._lib = _lib
        self.module = module
        self.params = [p for p in module.parameters() if p.requires_grad]
        dev = self.params[0].device
        sizes = [((p.numel() + 3) // 4) * 4 for p in self.params]     # keep every view 16-byte aligned
        n = sum(sizes)
        self.n = n
        self.flat_p = torch.zeros(n, device=dev)
        self.flat_g = torch.zeros(n, device=dev)
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        off = 0
        self.views = []
        for p, sz in zip(self.params, sizes):
            pv = self.flat_p[off:off + p.numel()].view_as(p)
            pv.copy_(p.data)
            p.data = pv
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)
            self.views.append((off, p.numel()))
            off += sz
        self.ema = self.flat_p.clone() if ema_decay > 0 else None
        self.state = torch.tensor([0.0, lr], device=dev)
        self.normsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.betas, self.eps, self.wd, self.max_norm, self.ema_decay = betas, eps, weight_decay, max_norm, ema_decay

    def zero_grad(self):
        self.flat_g.zero_()

    def set_lr(self, lr):
        self.state[1] = lr

    def step(self):
        lib, ptr, stream = self._lib.lib(), self._lib.ptr, self._lib.stream
        nq = None
        if self.max_norm > 0:
            self._lib.check(lib.ddg_grad_norm_sq(ptr(self.flat_g), self.n, ptr(self.normsq), stream()), 'grad_norm_sq')
            nq = self.normsq
        self._lib.check(lib.ddg_adam_ema_step(ptr(self.flat_p), ptr(self.flat_g), ptr(self.m), ptr(self.v), ptr(self.ema), self.n,
                                              ptr(self.state), ptr(nq), self.max_norm, self.betas[0], self.betas[1], self.eps,
                                              self.wd, self.ema_decay, stream()), 'adam_ema_step')
        # parameters changed behind autograd's back: tell the module so its fused inference engine re-packs its operands
        self.module._manual_version = getattr(self.module, '_manual_version', 0) + 1

    def ema_state_dict(self, module):
        out = {}
        for (name, p), (off, num) in zip([(n_, p_) for n_, p_ in module.named_parameters() if p_.requires_grad], self.views):
            out[name] = self.ema[off:off + num].view_as(p).clone()
        return out


class Trainer:
    """args needs: num_timesteps, beta_min, beta_max, use_geometric, nz, lr_g, lr_d, beta1, beta2, r1_gamma, lazy_reg,
    grad_clip_norm, ema_decay (names as in train_ddgan.py)."""

    def __init__(self, args, netG, netD, device, distributed=False, skip_discarded_g_backward=True, fused_optim=True):
        # ddgan.py:471-477 back-propagates errD_fake through the (un-detached) generator, and ddgan.py:489 zeroes those
        # generator gradients before anything reads them.  With skip_discarded_g_backward the D step evaluates G under no_grad
        # (fused inference plan): identical D gradients, identical parameter updates, none of the discarded work.
        self.skip_discarded_g_backward = skip_discarded_g_backward
        self.args = args
        self.netG, self.netD = netG, netD
        self.dev = device
        self.coeff = diffusion.DiffusionCoefficients(args, device)
        self.pos_coeff = diffusion.PosteriorCoefficients(args, device)
        betas = (getattr(args, 'beta1', 0.5), getattr(args, 'beta2', 0.9))
        # capturable=True keeps the Adam step counters on the device so that the whole step can be a CUDA graph
        cap = torch.device(device).type == 'cuda'
        self.fused_optim = fused_optim and cap
        wd = getattr(args, 'weight_decay', 0.0) or 0.0
        if self.fused_optim:
            # grad-norm -> clip -> Adam -> EMA as one flat-arena pass per network
            self.optD = FlatAdam(netD, args.lr_d, betas, weight_decay=wd, max_norm=args.grad_clip_norm)
            self.optG = FlatAdam(netG, args.lr_g, betas, weight_decay=wd, max_norm=args.grad_clip_norm,
                                 ema_decay=(getattr(args, 'ema_decay', 0.9999) if getattr(args, 'use_ema', True) else 0.0))
        else:
            self.optD = torch.optim.Adam(netD.parameters(), lr=args.lr_d, betas=betas, weight_decay=wd, capturable=cap, foreach=True)
            self.optG = torch.optim.Adam(netG.parameters(), lr=args.lr_g, betas=betas, weight_decay=wd, capturable=cap, foreach=True)
        self._graphs = None
        self.ema = None
        self.distributed = distributed and dist.is_initialized() and dist.get_world_size() > 1
        if self.distributed:
            broadcast_params(netG.parameters())
            broadcast_params(netD.parameters())
            if not self.fused_optim:
                self.arG = FlatGradAllReducer(netG.parameters())
                self.arD = FlatGradAllReducer(netD.parameters())
            elif self.optG.ema is not None:
                self.optG.ema.copy_(self.optG.flat_p)        # EMA starts from the broadcast weights
        if not self.fused_optim and getattr(args, 'use_ema', True):
            self.ema = EMA(netG, getattr(args, 'ema_decay', 0.9999))
        self.world = dist.get_world_size() if self.distributed else 1

    def _reduce_clip_step(self, opt, net, ar_name):
        if self.fused_optim:
            if self.distributed:
                dist.all_reduce(opt.flat_g)              # the arena IS the bucket: one collective, no copies
                opt.flat_g.div_(self.world)
            opt.step()
        else:
            if self.distributed:
                getattr(self, ar_name).allreduce()
            torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=self.args.grad_clip_norm)
            opt.step()

    def step(self, real_data, global_step, noise=None):
        """One iteration of ddgan.py:443-518.  `noise` (parity runs) = dict with t_d, n_xtp1_d, n_xt_d, z_d, n_post_d and the
        same with suffix _g; by default everything is drawn with torch's CUDA generator in the reference's order."""
        a, netG, netD = self.args, self.netG, self.netD
        nz = noise or {}
        B = real_data.size(0)
        # ---------------- D step ----------------
        for p in netD.parameters():
            p.requires_grad = True
        if self.fused_optim:
            self.optD.zero_grad()
        else:
            netD.zero_grad(set_to_none=False)   # grads stay allocated (outside any CUDA-graph pool) and are zeroed in place
        t = nz.get('t_d', None)
        if t is None:
            t = torch.randint(0, a.num_timesteps, (B,), device=self.dev)
        x_t, x_tp1 = diffusion.q_sample_pairs(self.coeff, real_data, t, noise_xt=nz.get('n_xt_d'), noise_xtp1=nz.get('n_xtp1_d'))
        x_t.requires_grad = True
        D_real = netD(x_t, t, x_tp1.detach()).view(-1)
        errD_real = F.softplus(-D_real).mean()
        do_r1 = (a.lazy_reg is None) or (global_step % a.lazy_reg == 0)
        errD_real.backward(retain_graph=do_r1)
        if do_r1:
            grad_real = torch.autograd.grad(outputs=D_real.sum(), inputs=x_t, create_graph=True)[0]
            grad_penalty = a.r1_gamma / 2 * (grad_real.view(B, -1).norm(2, dim=1) ** 2).mean()
            grad_penalty.backward()
        z = nz.get('z_d')
        if z is None:
            z = torch.randn(B, a.nz, device=self.dev)
        if self.skip_discarded_g_backward:
            with torch.no_grad():
                x_0_predict = netG(x_tp1.detach(), t, z)
                x_pos_sample = diffusion.sample_posterior(self.pos_coeff, x_0_predict, x_tp1.detach(), t, noise=nz.get('n_post_d'))
        else:
            x_0_predict = netG(x_tp1.detach(), t, z)
            x_pos_sample = diffusion.sample_posterior(self.pos_coeff, x_0_predict, x_tp1, t, noise=nz.get('n_post_d'))
        output = netD(x_pos_sample, t, x_tp1.detach()).view(-1)
        errD_fake = F.softplus(output).mean()
        errD_fake.backward()
        errD = errD_real.detach() + errD_fake.detach()
        self._reduce_clip_step(self.optD, netD, 'arD')
        # ---------------- G step ----------------
        for p in netD.parameters():
            p.requires_grad = False
        if self.fused_optim:
            self.optG.zero_grad()
        else:
            netG.zero_grad(set_to_none=False)
        t = nz.get('t_g', None)
        if t is None:
            t = torch.randint(0, a.num_timesteps, (B,), device=self.dev)
        x_t, x_tp1 = diffusion.q_sample_pairs(self.coeff, real_data, t, noise_xt=nz.get('n_xt_g'), noise_xtp1=nz.get('n_xtp1_g'))
        z = nz.get('z_g')
        if z is None:
            z = torch.randn(B, a.nz, device=self.dev)
        x_0_predict = netG(x_tp1.detach(), t, z)
        x_pos_sample = diffusion.sample_posterior(self.pos_coeff, x_0_predict, x_tp1, t, noise=nz.get('n_post_g'))
        output = netD(x_pos_sample, t, x_tp1.detach()).view(-1)
        errG = F.softplus(-output).mean()
        errG.backward()
        self._reduce_clip_step(self.optG, netG, 'arG')
        if self.ema is not None:
            self.ema.step()
        return errD, errG.detach()

    # ------------------------------------------------------------------------------------------------------------
    # whole-step CUDA graphs: the step is ~15 K kernel launches (launch-bound when issued from Python); captured once per
    # variant (with / without the lazy R1 double-backward) it replays as two graph launches per iteration.
    # ------------------------------------------------------------------------------------------------------------
    def capture(self, batch_shape, warmup=3, variants=('r1', 'plain'), share_pool=False):
        dev = self.dev
        self.real_static = torch.zeros(batch_shape, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(warmup):
                self.step(self.real_static, 0 if i == 0 else 1)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graphs = {}
        pool = None
        for key, gs in (('r1', 0), ('plain', 1)):
            if key not in variants:
                continue
            g = torch.cuda.CUDAGraph()
            # separate pools by default: the two variants are replayed in data-dependent order (1 : lazy_reg-1), which a shared
            # pool does not allow
            with torch.cuda.graph(g, pool=pool if share_pool else None):
                out = self.step(self.real_static, gs)
            pool = g.pool()
            self._graphs[key] = (g, out)
        return self

    def step_graphed(self, real_data, global_step):
        """Same semantics as step() (fresh randomness every replay through the graph-safe CUDA generator)."""
        if self._graphs is None:
            raise RuntimeError('call capture(batch_shape) first')
        a = self.args
        do_r1 = (a.lazy_reg is None) or (global_step % a.lazy_reg == 0)
        self.real_static.copy_(real_data, non_blocking=True)
        g, out = self._graphs['r1' if do_r1 else 'plain']
        g.replay()
        return out
