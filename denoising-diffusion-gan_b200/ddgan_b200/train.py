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
from . import ops


def broadcast_params(params, src=0, modules=()):
    """ddgan.py:30-33.  `modules`: engine-backed modules whose packed operands must be refreshed afterwards (the broadcast
    writes through p.data, which autograd's version counters do not see)."""
    for p in params:
        dist.broadcast(p.data, src=src)
    for m in modules:
        if hasattr(m, 'mark_dirty'):
            m.mark_dirty()


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
    """ema.py:45-55: ema = decay * ema + (1 - decay) * param, per parameter (multi-tensor ops).  Used with torch.optim.Adam
    (`fused_optim=False`); FlatAdam carries the EMA inside its own pass."""

    def __init__(self, model, ema_decay=0.999):
        self.decay = ema_decay
        self.model = model
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
    def load_state_dict(self, sd):
        for n, s in zip(self.names, self.shadow):
            if n in sd:
                s.copy_(sd[n].to(s.device))

    @torch.no_grad()
    def swap_parameters_with_ema(self, store_params_in_ema=True):
        """ema.py:57-79."""
        for p, s in zip(self.params, self.shadow):
            if store_params_in_ema:
                tmp = p.detach().clone()
                p.data.copy_(s)
                s.copy_(tmp)
            else:
                p.data.copy_(s)
        if hasattr(self.model, 'mark_dirty'):
            self.model.mark_dirty()


class FlatAdam:
    """clip_grad_norm_ + Adam (+ EMA) over flat parameter / gradient arenas: 3 launches per step instead of ~40 foreach
    kernels over hundreds of tensors.  The module's parameters and their .grad become views into the arenas, so the DP
    all-reduce runs on the gradient arena directly (no flatten / unflatten copies).

    Checkpoint surface (ddgan.py:545-569 saves optimizerG/optimizerD/emaG and EMA-swapped generator weights):
    `state_dict()` / `load_state_dict()` speak torch.optim.Adam's layout (per-parameter step / exp_avg / exp_avg_sq, one
    param group), `ema_state_dict()` / `load_ema_state_dict()` the reference EMA's {name: cpu tensor}, and
    `swap_parameters_with_ema()` is ema.py:57-79 on the arenas."""

    def __init__(self, module, lr, betas=(0.5, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.0, ema_decay=0.0, early=None):
        """early(name) -> bool: parameters whose gradient is final early in the backward pass are laid out first in the arenas
        (elements [0, n_early)), so that the data-parallel all-reduce of that contiguous range can overlap the rest of the
        backward.  Checkpoint I/O goes by name, so the order is invisible outside."""
        from . import _lib
        self._lib = _lib
        self.module = module
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if early is not None:
            named = [(n, p) for n, p in named if early(n)] + [(n, p) for n, p in named if not early(n)]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
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
        self.n_early = 0
        for (nme, p), sz in zip(named, sizes):
            pv = self.flat_p[off:off + p.numel()].view_as(p)
            pv.copy_(p.data)
            p.data = pv
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)
            self.views.append((off, p.numel()))
            off += sz
            if early is not None and early(nme):
                self.n_early = off
        self.ema = self.flat_p.clone() if ema_decay > 0 else None
        self.state = torch.tensor([0.0, lr], device=dev)
        self.normsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.betas, self.eps, self.wd, self.max_norm, self.ema_decay = (float(betas[0]), float(betas[1])), eps, float(weight_decay or 0.0), max_norm, ema_decay
        self.grad_scale = 1.0       # folded into the pass: 1/world after a sum all-reduce
        self.param_groups = [{'lr': lr}]   # what torch.optim.lr_scheduler reads and writes; synced to the device by step()
        self._lr_dev = lr
        self._mark()

    def _mark(self):
        if hasattr(self.module, 'mark_dirty'):
            self.module.mark_dirty()

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()

    def set_lr(self, lr):
        self.param_groups[0]['lr'] = lr

    def _sync_lr(self):
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_dev:
            self.state[1] = lr
            self._lr_dev = lr

    def step(self):
        lib, ptr, stream = self._lib.lib(), self._lib.ptr, self._lib.stream
        if not torch.cuda.is_current_stream_capturing():
            self._sync_lr()
        nq = None
        if self.max_norm > 0:
            self._lib.check(lib.ddg_grad_norm_sq(ptr(self.flat_g), self.n, ptr(self.normsq), stream()), 'grad_norm_sq')
            nq = self.normsq
        self._lib.check(lib.ddg_adam_ema_step(ptr(self.flat_p), ptr(self.flat_g), ptr(self.m), ptr(self.v), ptr(self.ema), self.n,
                                              ptr(self.state), ptr(nq), self.max_norm, self.betas[0], self.betas[1], self.eps,
                                              self.wd, self.ema_decay, self.grad_scale, stream()), 'adam_ema_step')
        # parameters changed behind autograd's back: the module's fused inference engines must re-pack their operands
        self._mark()

    # ---- checkpoint surface ----
    def _per_param(self, flat):
        return [flat[off:off + num].view_as(p) for (off, num), p in zip(self.views, self.params)]

    def state_dict(self):
        step = self.state[0].detach().clone()
        st = {}
        for i, (m, v) in enumerate(zip(self._per_param(self.m), self._per_param(self.v))):
            st[i] = {'step': step.clone(), 'exp_avg': m.clone(), 'exp_avg_sq': v.clone()}
        group = {'lr': float(self.param_groups[0]['lr']), 'betas': self.betas, 'eps': self.eps, 'weight_decay': self.wd,
                 'amsgrad': False, 'maximize': False, 'foreach': None, 'capturable': True, 'differentiable': False,
                 'fused': None, 'decoupled_weight_decay': False, 'params': list(range(len(self.params)))}
        return {'state': st, 'param_groups': [group]}

    @torch.no_grad()
    def load_state_dict(self, sd):
        g = sd['param_groups'][0]
        self.betas = (float(g['betas'][0]), float(g['betas'][1]))
        self.eps, self.wd = float(g['eps']), float(g['weight_decay'])
        self.set_lr(float(g['lr']))
        step = 0.0
        for i, (m, v) in enumerate(zip(self._per_param(self.m), self._per_param(self.v))):
            e = sd['state'].get(i, sd['state'].get(str(i)))
            if e is None:
                m.zero_(); v.zero_()
                continue
            m.copy_(e['exp_avg'].to(m.device)); v.copy_(e['exp_avg_sq'].to(v.device))
            step = max(step, float(e['step']))
        self.state[0] = step
        self._sync_lr()

    def ema_state_dict(self, module=None):
        """Reference layout (ema.py:81-83): {parameter name: CPU tensor}."""
        return {n: e.detach().cpu().clone() for n, e in zip(self.names, self._per_param(self.ema))}

    @torch.no_grad()
    def load_ema_state_dict(self, sd):
        for n, e in zip(self.names, self._per_param(self.ema)):
            if n in sd:
                e.copy_(sd[n].to(e.device))

    @torch.no_grad()
    def swap_parameters_with_ema(self, store_params_in_ema=True):
        """ema.py:57-79 on the arenas: parameters <-> EMA (or parameters <- EMA)."""
        if self.ema is None:
            return
        if store_params_in_ema:
            tmp = self.flat_p.clone()
            self.flat_p.copy_(self.ema)
            self.ema.copy_(tmp)
        else:
            self.flat_p.copy_(self.ema)
        self._mark()


def _arg(args, names, default):
    for n in names:
        v = getattr(args, n, None)
        if v is not None:
            return v
    return default


NOISE_KEYS = ('t', 'n_xtp1', 'n_xt', 'z', 'n_post')


class Trainer:
    """args (names as in train_ddgan.py): num_timesteps, beta_min, beta_max, use_geometric, nz, lr_g, lr_d, beta1_g / beta2_g /
    beta1_d / beta2_d (train_ddgan.py:86-89; the upstream `beta1` / `beta2` are the fallback), weight_decay_G / weight_decay_D
    (:73-84), r1_gamma, lazy_reg, grad_clip_norm, ema_decay, use_ema."""

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
        # ddgan.py:298-310: one Adam per network with its own betas and weight decay
        betas_d = (_arg(args, ('beta1_d', 'beta1'), 0.5), _arg(args, ('beta2_d', 'beta2'), 0.999))
        betas_g = (_arg(args, ('beta1_g', 'beta1'), 0.5), _arg(args, ('beta2_g', 'beta2'), 0.999))
        wd_d = _arg(args, ('weight_decay_D', 'weight_decay'), 0.0)
        wd_g = _arg(args, ('weight_decay_G', 'weight_decay'), 0.0)
        clip = _arg(args, ('grad_clip_norm',), 0.0)
        # capturable=True keeps the Adam step counters on the device so that the whole step can be a CUDA graph
        cap = torch.device(device).type == 'cuda'
        self.fused_optim = fused_optim and cap
        # conv / FIR outputs of the training graphs from ops.FRAME_POOL (no per-tensor border clears); DDG_TRAIN_FRAME_POOL=0 disables
        import os as _os
        self.frame_pool = cap and _os.environ.get('DDG_TRAIN_FRAME_POOL', '1') == '1'
        use_ema = bool(getattr(args, 'use_ema', True))
        ema_decay = _arg(args, ('ema_decay',), 0.9999) if use_ema else 0.0
        self.distributed = distributed and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.distributed else 1
        if self.distributed:
            broadcast_params(netG.parameters(), modules=(netG,))
            broadcast_params(netD.parameters(), modules=(netD,))
        if self.fused_optim:
            # grad-norm -> clip -> Adam -> EMA as one flat-arena pass per network
            self.optD = FlatAdam(netD, args.lr_d, betas_d, weight_decay=wd_d, max_norm=clip)
            early = None
            if self.distributed and hasattr(netG, 'cfg'):
                # up path + head: differentiated first (train_graph.generator_forward fires _grad_ready_hook when the backward
                # reaches the middle of the network); the AdaGN style / Dense_0 projections get their gradients at the very end
                from . import train_graph
                first = train_graph.up_first_idx(netG.cfg)

                def early(name, first=first):
                    if not name.startswith('all_modules.') or '.style.' in name or 'Dense_0' in name:
                        return False
                    return int(name.split('.')[1]) >= first
            self.optG = FlatAdam(netG, args.lr_g, betas_g, weight_decay=wd_g, max_norm=clip, ema_decay=ema_decay, early=early)
            # the all-reduce is a plain sum; the mean's 1/world rides in the optimiser pass
            self.optD.grad_scale = self.optG.grad_scale = 1.0 / self.world
        else:
            self.optD = torch.optim.Adam(netD.parameters(), lr=args.lr_d, betas=betas_d, weight_decay=wd_d, capturable=cap, foreach=True)
            self.optG = torch.optim.Adam(netG.parameters(), lr=args.lr_g, betas=betas_g, weight_decay=wd_g, capturable=cap, foreach=True)
            if self.distributed:
                self.arG = FlatGradAllReducer(netG.parameters())
                self.arD = FlatGradAllReducer(netD.parameters())
        self._graphs = None
        self._packs_frozen = False
        self.noise_static = None
        # data-parallel overlap (fused path): gradient all-reduces run on a side stream
        self._side = torch.cuda.Stream(device=device) if (self.distributed and self.fused_optim) else None
        self._g_armed = False
        self._d_pending = False
        if self._side is not None and self.optG.n_early > 0:
            from . import train_graph
            import weakref
            train_graph.GRAD_READY_HOOKS[netG] = weakref.WeakMethod(self._g_early_ready)   # no cycle netG -> Trainer -> netG
        self.ema = EMA(netG, ema_decay) if (not self.fused_optim and use_ema) else None

    # ---- checkpoint surface of the reference loop (ddgan.py:545-569) ----
    def ema_state_dict(self):
        return self.optG.ema_state_dict() if self.fused_optim else (self.ema.state_dict() if self.ema is not None else {})

    def swap_parameters_with_ema(self, store_params_in_ema=True):
        if self.fused_optim:
            self.optG.swap_parameters_with_ema(store_params_in_ema)
        elif self.ema is not None:
            self.ema.swap_parameters_with_ema(store_params_in_ema)

    def _g_early_ready(self):
        """Backward hook (train_graph.generator_forward): the up path and the head have been differentiated -> all-reduce their
        contiguous range of the gradient arena on the side stream while the down path is still in flight."""
        if not self._g_armed:
            return
        self._g_armed = False
        self._side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.optG.flat_g[:self.optG.n_early])

    def _d_reduce_async(self):
        """All D gradients are final after the fake-sample backward: reduce them on the side stream; the optimiser step waits for
        it just before D is next used (after the generator forward of the G step)."""
        self._side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.optD.flat_g)
        self._d_pending = True

    def _d_finish(self):
        if self._d_pending:
            torch.cuda.current_stream(self.dev).wait_stream(self._side)
            self.optD.step()
            self._d_pending = False

    def _reduce_clip_step(self, opt, net, ar_name):
        if self.fused_optim:
            if self.distributed:
                if opt is self.optG and self._side is not None and opt.n_early > 0:
                    dist.all_reduce(opt.flat_g[opt.n_early:])            # the rest; the early range is already in flight
                    torch.cuda.current_stream(self.dev).wait_stream(self._side)
                else:
                    dist.all_reduce(opt.flat_g)          # the arena IS the bucket: one collective (sum), no copies, no div
            opt.step()
        else:
            if self.distributed:
                getattr(self, ar_name).allreduce()
            torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=self.args.grad_clip_norm)
            opt.step()

    def draw_noise(self, B, shape, out=None):
        """Every random tensor of one iteration, in the reference's draw order (ddgan.py:450,122,112,470,164 and the same in
        the G step).  `out`: dict of static buffers to refill in place (CUDA-graph replays read them)."""
        a, dev = self.args, self.dev
        nz = out if out is not None else {}
        for sfx in ('_d', '_g'):
            if out is None:
                nz['t' + sfx] = torch.randint(0, a.num_timesteps, (B,), device=dev)
                nz['n_xtp1' + sfx] = torch.randn(shape, device=dev)
                nz['n_xt' + sfx] = torch.randn(shape, device=dev)
                nz['z' + sfx] = torch.randn(B, a.nz, device=dev)
                nz['n_post' + sfx] = torch.randn(shape, device=dev)
            else:
                nz['t' + sfx].random_(0, a.num_timesteps)
                for k in ('n_xtp1', 'n_xt', 'z', 'n_post'):
                    nz[k + sfx].normal_()
        return nz

    def step(self, real_data, global_step, noise=None):
        """One iteration of ddgan.py:443-518.  `noise` (parity runs / graph replays) = dict with t_d, n_xtp1_d, n_xt_d, z_d,
        n_post_d and the same with suffix _g; by default everything is drawn with torch's CUDA generator in the reference's
        order."""
        from . import train_graph
        # with the flat arenas every parameter owns its .grad for good: let the wgrad / bias-sum kernels accumulate into it
        train_graph.ACCUM['on'] = self.fused_optim
        if self.frame_pool:
            ops.FRAME_POOL.begin_step()      # activations of this step come from the persistent frame-preserving pool
        try:
            out = self._step(real_data, global_step, noise)
        finally:
            train_graph.ACCUM['on'] = False
            ops.FRAME_POOL.end_step()
        if not self._packs_frozen:
            # the first step recorded every weight pack of both networks; from now on one launch per network forward
            train_graph.freeze_packs(self.netG)
            train_graph.freeze_packs(self.netD)
            self._packs_frozen = True
        return out

    def _step(self, real_data, global_step, noise=None):
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
        if self._side is not None:
            self._d_reduce_async()       # overlaps the generator forward of the G step; finished by _d_finish() below
        else:
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
        self._d_finish()                 # D's update (ddgan.py:485) must land before D scores the G-step sample
        output = netD(x_pos_sample, t, x_tp1.detach()).view(-1)
        errG = F.softplus(-output).mean()
        self._g_armed = self._side is not None
        errG.backward()
        self._g_armed = False
        self._reduce_clip_step(self.optG, netG, 'arG')
        if self.ema is not None:
            self.ema.step()
        return errD, errG.detach()

    # ------------------------------------------------------------------------------------------------------------
    # whole-step CUDA graphs: the step is thousands of kernel launches (launch-bound when issued from Python); captured once
    # per variant (with / without the lazy R1 double-backward) it replays as one graph launch per iteration.
    # ------------------------------------------------------------------------------------------------------------
    def _snapshot(self):
        """Everything a training step mutates: parameters, optimiser moments / step counters, EMA."""
        if self.fused_optim:
            return [(t, t.clone()) for o in (self.optD, self.optG) for t in (o.flat_p, o.m, o.v, o.state, o.ema) if t is not None]
        snap = [(p.data, p.detach().clone()) for net in (self.netD, self.netG) for p in net.parameters()]
        for o in (self.optD, self.optG):
            for st in o.state.values():
                snap += [(v, v.clone()) for v in st.values() if torch.is_tensor(v)]
        if self.ema is not None:
            snap += [(s, s.clone()) for s in self.ema.shadow]
        return snap

    def capture(self, batch_shape, warmup=3, variants=('r1', 'plain'), share_pool=False, static_noise=True):
        """Captures the step once per variant.  The warm-up iterations needed before capture (lazy initialisation, allocator
        pools) run real optimiser steps on a dummy batch, so the training state is snapshotted first and restored in place
        afterwards: capturing on a pretrained / resumed model leaves weights, Adam moments, step counters and EMA untouched.

        static_noise: the step's random tensors live in static buffers (`self.noise_static`) that `step_graphed` refills (or
        copies caller-provided noise into) before each replay -- the graph itself is then a pure function of
        (real, noise, state), which is what the parity test replays against the eager step."""
        dev = self.dev
        B = batch_shape[0]
        self.real_static = torch.zeros(batch_shape, device=dev)
        self.noise_static = self.draw_noise(B, tuple(batch_shape)) if static_noise else None
        snap = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(warmup):
                self.step(self.real_static, 0 if i == 0 else 1, noise=self.noise_static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for dst, src in snap:
                dst.copy_(src)
        for net in (self.netG, self.netD):
            if hasattr(net, 'mark_dirty'):
                net.mark_dirty()
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
                out = self.step(self.real_static, gs, noise=self.noise_static)
            pool = g.pool()
            self._graphs[key] = (g, out)
        return self

    def step_graphed(self, real_data, global_step, noise=None):
        """Same semantics as step(): fresh randomness every replay (static noise buffers refilled here, or -- without
        static_noise -- drawn inside the graph by the graph-safe CUDA generator); `noise` injects the draws instead."""
        if self._graphs is None:
            raise RuntimeError('call capture(batch_shape) first')
        a = self.args
        do_r1 = (a.lazy_reg is None) or (global_step % a.lazy_reg == 0)
        self.real_static.copy_(real_data, non_blocking=True)
        if self.noise_static is not None:
            if noise is not None:
                for k, buf in self.noise_static.items():
                    buf.copy_(noise[k])
            else:
                self.draw_noise(real_data.size(0), None, out=self.noise_static)
        elif noise is not None:
            raise RuntimeError('noise injection needs capture(static_noise=True)')
        if self.fused_optim:
            self.optD._sync_lr(); self.optG._sync_lr()
        g, out = self._graphs['r1' if do_r1 else 'plain']
        g.replay()
        # the replay updated both networks through raw pointers: their cached inference engines are stale
        for net in (self.netG, self.netD):
            if hasattr(net, 'mark_dirty'):
                net.mark_dirty()
        return out
