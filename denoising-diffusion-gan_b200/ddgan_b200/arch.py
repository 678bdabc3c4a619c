"""Architecture descriptions: the module sequence and the state_dict names/shapes of the reference models.

Checkpoint compatibility is part of the drop-in contract (test_ddgan.py:162 loads with strict=True), so parameter
names and shapes follow NCSNpp.__init__ (score_sde/models/ncsnpp_generator_adagn.py:93-277) and
Discriminator_small/large.__init__ (score_sde/models/discriminator.py:96-132, 170-203) exactly.
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace

GEN_DEFAULTS = dict(
    image_size=32, num_channels=3, num_channels_dae=128, ch_mult=(1, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,),
    dropout=0.0, resamp_with_conv=True, conditional=True, fir=True, fir_kernel=[1, 3, 3, 1], skip_rescale=True,
    resblock_type='biggan', progressive='none', progressive_input='residual', progressive_combine='sum',
    embedding_type='positional', fourier_scale=16.0, not_use_tanh=False, z_emb_dim=256, nz=100, n_mlp=4, centered=True,
    t_emb_dim=256, ngf=64, num_timesteps=4, beta_min=0.1, beta_max=20.0, use_geometric=False)


def make_config(**over):
    """CIFAR-10 configuration of the reference README (readme.md:31-37) with overrides."""
    d = dict(GEN_DEFAULTS)
    d.update(over)
    return SimpleNamespace(**d)


def normalize_config(cfg):
    """Accept argparse.Namespace / SimpleNamespace / dict; fill attributes NCSNpp.__init__ reads (:63-91,176,272-274)."""
    d = dict(GEN_DEFAULTS)
    src = cfg if isinstance(cfg, dict) else vars(cfg)
    d.update({k: v for k, v in src.items() if v is not None or k not in d})
    d['ch_mult'] = tuple(d['ch_mult'])
    d['attn_resolutions'] = tuple(d['attn_resolutions'])
    d['fir_kernel'] = list(d['fir_kernel'])
    for k in ('resblock_type', 'progressive', 'progressive_input', 'embedding_type', 'progressive_combine'):
        d[k] = str(d[k]).lower()
    return SimpleNamespace(**d)


def ncsnpp_modules(cfg):
    """Ordered `all_modules` entries as dicts: kind in {linear, conv3, res, attn, pyrdown, gn}."""
    nf = cfg.num_channels_dae
    nres = len(cfg.ch_mult)
    res = [cfg.image_size // 2 ** i for i in range(nres)]
    mods = []

    def add(kind, **kw):
        mods.append(dict(kind=kind, idx=len(mods), **kw))

    if cfg.conditional:
        add('linear', i=nf, o=4 * nf)
        add('linear', i=4 * nf, o=4 * nf)
    add('conv3', i=cfg.num_channels, o=nf)
    hs_c = [nf]
    in_ch = nf
    pyr_ch = cfg.num_channels
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            out_ch = nf * cfg.ch_mult[lvl]
            add('res', i=in_ch, o=out_ch, up=False, down=False, res=res[lvl])
            in_ch = out_ch
            if res[lvl] in cfg.attn_resolutions:
                add('attn', c=in_ch, res=res[lvl])
            hs_c.append(in_ch)
        if lvl != nres - 1:
            add('res', i=in_ch, o=in_ch, up=False, down=True, res=res[lvl])
            if cfg.progressive_input == 'residual':
                add('pyrdown', i=pyr_ch, o=in_ch, res=res[lvl])
                pyr_ch = in_ch
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    add('res', i=in_ch, o=in_ch, up=False, down=False, res=res[-1])
    add('attn', c=in_ch, res=res[-1])
    add('res', i=in_ch, o=in_ch, up=False, down=False, res=res[-1])
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            out_ch = nf * cfg.ch_mult[lvl]
            add('res', i=in_ch + hs_c.pop(), o=out_ch, up=False, down=False, res=res[lvl])
            in_ch = out_ch
        if res[lvl] in cfg.attn_resolutions:
            add('attn', c=in_ch, res=res[lvl])
        if lvl != 0:
            add('res', i=in_ch, o=in_ch, up=True, down=False, res=res[lvl])
    assert not hs_c
    add('gn', c=in_ch)
    add('conv3', i=in_ch, o=cfg.num_channels)
    return mods


def ncsnpp_param_shapes(cfg) -> 'OrderedDict[str, tuple]':
    nf, zd = cfg.num_channels_dae, cfg.z_emb_dim
    shapes = OrderedDict()
    for m in ncsnpp_modules(cfg):
        p = f"all_modules.{m['idx']}."
        k = m['kind']
        if k == 'linear':
            shapes[p + 'weight'] = (m['o'], m['i']); shapes[p + 'bias'] = (m['o'],)
        elif k == 'conv3':
            shapes[p + 'weight'] = (m['o'], m['i'], 3, 3); shapes[p + 'bias'] = (m['o'],)
        elif k == 'res':
            i, o = m['i'], m['o']
            shapes[p + 'GroupNorm_0.style.weight'] = (2 * i, zd); shapes[p + 'GroupNorm_0.style.bias'] = (2 * i,)
            shapes[p + 'Conv_0.weight'] = (o, i, 3, 3); shapes[p + 'Conv_0.bias'] = (o,)
            shapes[p + 'Dense_0.weight'] = (o, 4 * nf); shapes[p + 'Dense_0.bias'] = (o,)
            shapes[p + 'GroupNorm_1.style.weight'] = (2 * o, zd); shapes[p + 'GroupNorm_1.style.bias'] = (2 * o,)
            shapes[p + 'Conv_1.weight'] = (o, o, 3, 3); shapes[p + 'Conv_1.bias'] = (o,)
            if i != o or m['up'] or m['down']:
                shapes[p + 'Conv_2.weight'] = (o, i, 1, 1); shapes[p + 'Conv_2.bias'] = (o,)
        elif k == 'attn':
            c = m['c']
            shapes[p + 'GroupNorm_0.weight'] = (c,); shapes[p + 'GroupNorm_0.bias'] = (c,)
            for j in range(4):
                shapes[p + f'NIN_{j}.W'] = (c, c); shapes[p + f'NIN_{j}.b'] = (c,)
        elif k == 'pyrdown':
            shapes[p + 'Conv2d_0.weight'] = (m['o'], m['i'], 3, 3); shapes[p + 'Conv2d_0.bias'] = (m['o'],)
        elif k == 'gn':
            shapes[p + 'weight'] = (m['c'],); shapes[p + 'bias'] = (m['c'],)
    shapes['z_transform.1.weight'] = (zd, cfg.nz); shapes['z_transform.1.bias'] = (zd,)
    for i in range(cfg.n_mlp):
        shapes[f'z_transform.{3 + 2 * i}.weight'] = (zd, zd); shapes[f'z_transform.{3 + 2 * i}.bias'] = (zd,)
    return shapes


def discriminator_blocks(ngf: int, large: bool):
    """(in, out, downsample) of conv1..convK (discriminator.py:112-123 small, :184-194 large)."""
    if large:
        ch = [(2, 4, True), (4, 8, True), (8, 8, True), (8, 8, True), (8, 8, True), (8, 8, True)]
    else:
        ch = [(2, 2, False), (2, 4, True), (4, 8, True), (8, 8, True)]
    return [(ngf * a, ngf * b, d) for a, b, d in ch]


def discriminator_param_shapes(nc: int, ngf: int, t_emb_dim: int, large: bool = False) -> 'OrderedDict[str, tuple]':
    s = OrderedDict()
    s['t_embed.main.0.weight'] = (t_emb_dim, t_emb_dim); s['t_embed.main.0.bias'] = (t_emb_dim,)
    s['t_embed.main.2.weight'] = (t_emb_dim, t_emb_dim); s['t_embed.main.2.bias'] = (t_emb_dim,)
    s['start_conv.weight'] = (2 * ngf, nc, 1, 1); s['start_conv.bias'] = (2 * ngf,)
    for i, (a, b, _) in enumerate(discriminator_blocks(ngf, large)):
        p = f'conv{i + 1}.'
        s[p + 'conv1.0.weight'] = (b, a, 3, 3); s[p + 'conv1.0.bias'] = (b,)
        s[p + 'conv2.0.weight'] = (b, b, 3, 3); s[p + 'conv2.0.bias'] = (b,)
        s[p + 'dense_t1.weight'] = (b, t_emb_dim); s[p + 'dense_t1.bias'] = (b,)
        s[p + 'skip.0.weight'] = (b, a, 1, 1)
    s['final_conv.weight'] = (8 * ngf, 8 * ngf + 1, 3, 3); s['final_conv.bias'] = (8 * ngf,)
    s['end_linear.weight'] = (1, 8 * ngf); s['end_linear.bias'] = (1,)
    return s
