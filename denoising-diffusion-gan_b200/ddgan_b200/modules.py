"""nn.Module shells with the reference's state_dict layout; the math runs in the fused engines (engine.py)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import arch
from .engine import DiscriminatorEngine, GeneratorEngine


class _Box(nn.Module):
    """Anonymous container: only exists so that state_dict keys read 'all_modules.3.Conv_0.weight' etc."""


def _register_tree(root: nn.Module, shapes):
    params = {}
    for name, shp in shapes.items():
        parts = name.split('.')
        mod = root
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, _Box())
            mod = getattr(mod, p)
        prm = nn.Parameter(torch.zeros(shp))
        mod.register_parameter(parts[-1], prm)
        params[name] = prm
    return params


def _uniform_(t, scale, denom):
    bound = math.sqrt(3.0 * scale / max(1.0, denom))
    with torch.no_grad():
        return t.uniform_(-bound, bound)


def _fans(shape, in_axis=1, out_axis=0):
    rf = 1
    for d in shape:
        rf *= d
    rf = rf / shape[in_axis] / shape[out_axis]
    return shape[in_axis] * rf, shape[out_axis] * rf


def _default_init_(t, scale=1.0):
    """layers.py:66-105: variance_scaling(scale, 'fan_avg', 'uniform'); scale 0 -> 1e-10."""
    scale = 1e-10 if scale == 0 else scale
    fi, fo = _fans(t.shape)
    return _uniform_(t, scale, (fi + fo) / 2)


def _dense_layer_init_(t, scale=1.0):
    """dense_layer.py:23-64: kaiming_uniform_ with mode 'fan_avg', which that file resolves to fan_out (:33)."""
    scale = 1e-10 if scale == 0 else scale
    fo = t.shape[0] * (t[0][0].numel() if t.dim() > 2 else 1)
    return _uniform_(t, scale, fo)


class _EngineModule(nn.Module):
    """Caches one fused inference engine per (batch, device, precision) and re-packs its operands when the parameters may
    have changed.

    Staleness rule (no hashing, no collisions):
      * in train() mode every engine call re-packs -- the weights are being optimised, possibly behind autograd's back
        (CUDA-graph replays, fused optimisers, `p.data.copy_` as in the reference's ema.py:70-79 / ddgan.py:30-33);
      * in eval() mode the engine is reused while `(epoch, per-parameter _version, per-parameter data_ptr)` is unchanged;
        `mark_dirty()` bumps the epoch and must be called by anything that writes parameters without going through a
        version-counted in-place op (train.FlatAdam, Trainer.step_graphed, train.EMA, train.broadcast_params do)."""

    def __init__(self):
        super().__init__()
        self._engines = {}
        self._engine_ptrs = None
        self._epoch = 0
        self.precision = 3  # 3 = BF16x3 (fp32 parity), 1 = BF16

    def mark_dirty(self):
        self._epoch += 1

    def _param_key(self):
        ps = list(self.parameters())
        return (self._epoch, tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))

    def _get_engine(self, batch, device, build):
        """build(params) -> engine reading the module's own parameter storage (no copies per update).  If that storage moved
        (FlatAdam re-homes the parameters into its arena, .to(...)), every cached engine is dropped and rebuilt."""
        k = self._param_key()
        if self._engines and self._engine_ptrs != k[2]:
            self._engines = {}
        key = (batch, str(device), self.precision)
        ent = self._engines.get(key)
        if ent is None:
            self._engine_ptrs = k[2]
            ent = {'eng': build({n: p for n, p in self.named_parameters()}), 'key': None}
            self._engines[key] = ent
        if self.training or ent['key'] != k:
            ent['eng'].refresh()
            ent['key'] = k
        return ent['eng']

    def load_state_dict(self, *a, **kw):
        self.mark_dirty()
        return super().load_state_dict(*a, **kw)

    def _apply(self, fn, *a, **kw):  # .to()/.cuda() invalidate engines
        self._engines = {}
        self._epoch += 1
        return super()._apply(fn, *a, **kw)


class NCSNpp(_EngineModule):
    """NCSN++ generator with AdaGN (reference: score_sde/models/ncsnpp_generator_adagn.py:59-431)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        cfg = arch.normalize_config(config)
        self.cfg = cfg
        self.not_use_tanh = cfg.not_use_tanh
        self.nf = cfg.num_channels_dae
        self.z_emb_dim = cfg.z_emb_dim
        shapes = arch.ncsnpp_param_shapes(cfg)
        prm = _register_tree(self, shapes)
        for name, p in prm.items():
            if name.endswith('.bias') or name.endswith('.b'):
                nn.init.zeros_(p)
                if name.endswith('style.bias'):
                    p.data[: p.shape[0] // 2] = 1  # layerspp.py:53-54
            elif name.endswith('GroupNorm_0.weight') or (p.dim() == 1 and name.endswith('.weight')):
                nn.init.ones_(p)
            elif 'style.weight' in name or name.startswith('z_transform'):
                _dense_layer_init_(p.data)
            elif name.endswith('.W'):
                _default_init_(p.data, 0.0 if name.endswith('NIN_3.W') else 0.1)
            else:
                zero_scale = name.endswith('Conv_1.weight') or name == f'all_modules.{len(arch.ncsnpp_modules(cfg)) - 1}.weight'
                _default_init_(p.data, 0.0 if zero_scale else 1.0)

    def forward(self, x, time_cond, z):
        dropout_live = self.training and float(self.cfg.dropout) > 0     # the fused inference plan has no dropout masks
        if dropout_live or (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            from . import train_graph
            return train_graph.generator_forward(self, x, time_cond, z)
        eng = self._get_engine(x.shape[0], x.device,
                               lambda prm: GeneratorEngine(self.cfg, x.shape[0], x.device, self.precision, params=prm))
        return eng.forward(x, time_cond, z).clone()


class _Discriminator(_EngineModule):
    large = False

    def __init__(self, nc=3, ngf=64, t_emb_dim=128, act=None):
        super().__init__()
        if act is not None and not (isinstance(act, nn.LeakyReLU) and abs(act.negative_slope - 0.2) < 1e-12):
            raise NotImplementedError('the fused discriminator implements act = LeakyReLU(0.2) (ddgan.py:281,289)')
        self.nc, self.ngf, self.t_emb_dim = nc, ngf, t_emb_dim
        self.act = nn.LeakyReLU(0.2)
        self.stddev_group, self.stddev_feat = 4, 1
        prm = _register_tree(self, arch.discriminator_param_shapes(nc, ngf, t_emb_dim, self.large))
        for name, p in prm.items():
            if name.endswith('.bias'):
                nn.init.zeros_(p)
            else:
                zero = name.endswith('conv2.0.weight') or (name == 'final_conv.weight' and not self.large)
                _dense_layer_init_(p.data, 0.0 if zero else 1.0)

    def forward(self, x, t, x_t):
        if torch.is_grad_enabled() and (x.requires_grad or x_t.requires_grad or any(p.requires_grad for p in self.parameters())):
            from . import train_graph
            return train_graph.discriminator_forward(self, x, t, x_t)
        S = x.shape[-1]
        eng = self._get_engine(x.shape[0], x.device,
                               lambda prm: DiscriminatorEngine(self.nc, self.ngf, self.t_emb_dim, S, x.shape[0], large=self.large,
                                                               device=x.device, precision=self.precision, params=prm))
        return eng.forward(x, t, x_t).clone()


class Discriminator_small(_Discriminator):
    """discriminator.py:96-167"""
    large = False


class Discriminator_large(_Discriminator):
    """discriminator.py:170-238"""
    large = True

    def __init__(self, nc=1, ngf=32, t_emb_dim=128, act=None):
        super().__init__(nc, ngf, t_emb_dim, act)
