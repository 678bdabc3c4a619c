"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    TORCH_EXTENSIONS_DIR=/tmp/torch_ext TORCH_CUDA_ARCH_LIST=10.0a python tests/golden/make_golden.py

The reference has no tests / golden vectors of its own, so the oracle is pinned against
outputs of the reference's own CPU path (upfirdn2d_native + torch CPU ops).  All inputs and weights are
regenerated from seeds by `oracle.ddgan_oracle.randomize_params` / torch.Generator, so the fixtures hold only
outputs (small) plus the reference's state_dict name/shape lists.
"""
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
os.environ.setdefault('TORCH_EXTENSIONS_DIR', '/tmp/torch_ext')
os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
sys.modules.setdefault('nibabel', types.ModuleType('nibabel'))

from oracle import ddgan_oracle as O  # noqa: E402

import score_sde.op as ref_op  # noqa: E402
from score_sde.op.upfirdn2d import upfirdn2d_native  # noqa: E402
from score_sde.models import up_or_down_sampling as ref_uds  # noqa: E402
from score_sde.models.ncsnpp_generator_adagn import NCSNpp  # noqa: E402
from score_sde.models.discriminator import Discriminator_small, Discriminator_large  # noqa: E402
import test_ddgan as ref_test  # noqa: E402
import torch.nn.functional as F  # noqa: E402

try:
    import ddgan as ref_ddgan
except Exception as e:  # pragma: no cover
    print('import ddgan failed:', e)
    ref_ddgan = None


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def ns(cfg):
    d = dict(vars(cfg))
    d['ch_mult'] = list(d['ch_mult'])
    d['attn_resolutions'] = list(d['attn_resolutions'])
    return Namespace(**d)


def main():
    out = {}
    torch.set_num_threads(8)

    # ---- KATs ----
    out['setup_kernel_1331'] = torch.from_numpy(ref_uds._setup_kernel([1, 3, 3, 1]))
    out['upsample_kat'] = ref_uds.upsample_2d(torch.tensor([[1., 2.], [3., 4.]]).view(1, 1, 2, 2), [1, 3, 3, 1], factor=2)
    out['downsample_kat'] = ref_uds.downsample_2d(torch.arange(16.).view(1, 1, 4, 4), [1, 3, 3, 1], factor=2)

    # ---- upfirdn2d_native sweep ----
    cases = []
    k4 = torch.from_numpy(ref_uds._setup_kernel([1, 3, 3, 1]))
    k3 = seeded((3, 3), 11)
    k2 = seeded((2, 2), 12)
    k5 = seeded((5, 5), 13)
    specs = [
        ('up2', (2, 3, 8, 8), k4 * 4, 2, 1, (2, 1)),
        ('down2', (2, 3, 8, 8), k4, 1, 2, (1, 1)),
        ('pad22', (2, 3, 8, 8), k4, 1, 1, (2, 2)),
        ('pad11_adj', (2, 3, 9, 9), k4, 1, 1, (1, 1)),
        ('k3_asym', (1, 2, 7, 5), k3, 1, 1, (1, 1)),
        ('k2_up2', (1, 2, 5, 6), k2, 2, 1, (1, 0)),
        ('k5_down2', (1, 2, 11, 9), k5, 1, 2, (2, 2)),
        ('up2_down2', (1, 1, 6, 6), k4, 2, 2, (2, 1)),
        ('negpad', (1, 2, 8, 8), k3, 1, 1, (-1, 2)),
        ('up3', (1, 1, 4, 4), k5, 3, 1, (3, 1)),
        ('odd_hw_down2', (1, 2, 7, 9), k4, 1, 2, (1, 1)),
    ]
    for i, (name, shp, k, up, down, pad) in enumerate(specs):
        x = seeded(shp, 100 + i)
        y = upfirdn2d_native(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        cases.append(dict(name=name, shape=shp, seed=100 + i, kernel=k.clone(), up=up, down=down, pad=pad, out=y))
    out['upfirdn2d_cases'] = cases

    # conv_downsample_2d
    x = seeded((2, 3, 8, 8), 130); w = seeded((5, 3, 3, 3), 131, 0.2)
    out['conv_downsample'] = ref_uds.conv_downsample_2d(x, w, [1, 3, 3, 1])

    # fused_leaky_relu (CPU branch: slope fixed 0.2)
    x = seeded((2, 5, 4, 4), 140); b = seeded((5,), 141)
    out['fused_lrelu'] = ref_op.fused_leaky_relu(x, b, 0.2, 2 ** 0.5)

    # ---- schedules ----
    for T in (2, 4):
        a = Namespace(num_timesteps=T, beta_min=0.1, beta_max=20.0, use_geometric=False)
        pc = ref_test.Posterior_Coefficients(a, 'cpu')
        sig, a_s, betas = ref_test.get_sigma_schedule(a, 'cpu')
        d = dict(sigmas=sig, a_s=a_s, betas=betas, coef1=pc.posterior_mean_coef1, coef2=pc.posterior_mean_coef2,
                 logvar=pc.posterior_log_variance_clipped)
        if ref_ddgan is not None:
            co = ref_ddgan.DiffusionCoefficients(a, 'cpu')
            d.update(a_s_cum=co.a_s_cum, sigmas_cum=co.sigmas_cum)
        out[f'sched_T{T}'] = d

    # ---- NCSN++ tiny ----
    cfg = O.tiny_config()
    net = NCSNpp(ns(cfg)).eval()
    ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    out['ncsnpp_tiny_shapes'] = ref_shapes
    sd = O.randomize_params(ref_shapes, seed=7)
    net.load_state_dict(sd, strict=True)
    B = 3
    x = seeded((B, 3, 16, 16), 200); z = seeded((B, cfg.nz), 201)
    t = torch.tensor([0, 3, 1])
    with torch.no_grad():
        out['ncsnpp_tiny_out'] = net(x, t, z)

    # cifar config: shapes only (+ param count)
    cfgc = O.cifar10_config()
    netc = NCSNpp(ns(cfgc))
    out['ncsnpp_cifar_shapes'] = {k: tuple(v.shape) for k, v in netc.state_dict().items()}
    # one cifar forward at B=2 (output + a checksum only)
    sdc = O.randomize_params(out['ncsnpp_cifar_shapes'], seed=8)
    netc.load_state_dict(sdc, strict=True); netc.eval()
    xc = seeded((2, 3, 32, 32), 210); zc = seeded((2, 100), 211); tc = torch.tensor([3, 0])
    with torch.no_grad():
        out['ncsnpp_cifar_out'] = netc(xc, tc, zc)
    del netc, sdc

    # ---- discriminators (tiny ngf) ----
    ds = Discriminator_small(nc=6, ngf=16, t_emb_dim=32).eval()
    out['dsmall_shapes'] = {k: tuple(v.shape) for k, v in ds.state_dict().items()}
    sdd = O.randomize_params(out['dsmall_shapes'], seed=9)
    ds.load_state_dict(sdd, strict=True)
    xd = seeded((4, 3, 32, 32), 220); xtd = seeded((4, 3, 32, 32), 221); td = torch.tensor([0, 1, 2, 3])
    with torch.no_grad():
        out['dsmall_out'] = ds(xd, td, xtd)
    dl = Discriminator_large(nc=6, ngf=8, t_emb_dim=32).eval()
    out['dlarge_shapes'] = {k: tuple(v.shape) for k, v in dl.state_dict().items()}
    sdl = O.randomize_params(out['dlarge_shapes'], seed=10)
    dl.load_state_dict(sdl, strict=True)
    xl = seeded((4, 3, 256, 256), 230); xtl = seeded((4, 3, 256, 256), 231)
    with torch.no_grad():
        out['dlarge_out'] = dl(xl, td, xtl)
    out['dsmall_cifar_shapes'] = {k: tuple(v.shape) for k, v in Discriminator_small(nc=6, ngf=64, t_emb_dim=256).state_dict().items()}

    # ---- sampler (seeded torch.randn order: z then posterior noise) ----
    a = Namespace(num_timesteps=4, beta_min=0.1, beta_max=20.0, use_geometric=False, nz=cfg.nz)
    pc = ref_test.Posterior_Coefficients(a, 'cpu')
    torch.manual_seed(1024)
    x_init = torch.randn(B, 3, 16, 16)
    out['sample_tiny'] = ref_test.sample_from_model(pc, net, 4, x_init, None, a)

    # ---- train step losses + gradients (tiny G + tiny D_small at 16px is invalid: D_small wants 32px) ----
    if ref_ddgan is not None:
        cfg32 = O.tiny_config(image_size=32, attn_resolutions=(16,))
        netg = NCSNpp(ns(cfg32))
        shp_g = {k: tuple(v.shape) for k, v in netg.state_dict().items()}
        out['ncsnpp_tiny32_shapes'] = shp_g
        netg.load_state_dict(O.randomize_params(shp_g, seed=21), strict=True)
        netd = Discriminator_small(nc=6, ngf=16, t_emb_dim=32)
        netd.load_state_dict(O.randomize_params(out['dsmall_shapes'], seed=22), strict=True)
        a = Namespace(num_timesteps=4, beta_min=0.1, beta_max=20.0, use_geometric=False, nz=cfg32.nz)
        coeff = ref_ddgan.DiffusionCoefficients(a, 'cpu'); pc = ref_ddgan.PosteriorCoefficients(a, 'cpu')
        Bt = 4
        real = torch.tanh(seeded((Bt, 3, 32, 32), 300))
        t = torch.tensor([0, 1, 2, 3])
        n_xtp1, n_xt, n_post = seeded((Bt, 3, 32, 32), 301), seeded((Bt, 3, 32, 32), 302), seeded((Bt, 3, 32, 32), 303)
        z = seeded((Bt, cfg32.nz), 304)
        # reproduce ddgan.py:449-477 with the randn draws replaced by the seeded tensors above
        x_t = coeff.a_s_cum[t].view(-1, 1, 1, 1) * real + coeff.sigmas_cum[t].view(-1, 1, 1, 1) * n_xt
        x_tp1 = coeff.a_s[t + 1].view(-1, 1, 1, 1) * x_t + coeff.sigmas[t + 1].view(-1, 1, 1, 1) * n_xtp1
        # cross-check against the reference's own q_sample with noise passed explicitly
        assert torch.allclose(x_t, ref_ddgan.q_sample(coeff, real, t, noise=n_xt))
        x_t.requires_grad = True
        netd.zero_grad(); netg.zero_grad()
        D_real = netd(x_t, t, x_tp1.detach()).view(-1)
        errD_real = F.softplus(-D_real).mean()
        errD_real.backward(retain_graph=True)
        grad_real = torch.autograd.grad(outputs=D_real.sum(), inputs=x_t, create_graph=True)[0]
        gp = (grad_real.view(grad_real.size(0), -1).norm(2, dim=1) ** 2).mean()
        gp = 0.5 / 2 * gp
        gp.backward()
        x0p = netg(x_tp1.detach(), t, z)
        mean = pc.posterior_mean_coef1[t].view(-1, 1, 1, 1) * x0p + pc.posterior_mean_coef2[t].view(-1, 1, 1, 1) * x_tp1
        x_pos = mean + (1 - (t == 0).float()).view(-1, 1, 1, 1) * torch.exp(0.5 * pc.posterior_log_variance_clipped[t].view(-1, 1, 1, 1)) * n_post
        output = netd(x_pos, t, x_tp1.detach()).view(-1)
        errD_fake = F.softplus(output).mean()
        errD_fake.backward()
        out['train_tiny'] = dict(
            errD_real=errD_real.detach(), gp=gp.detach(), errD_fake=errD_fake.detach(), r1_gamma=0.5,
            gradD={k: v.grad.clone() for k, v in netd.named_parameters()},
            gradG_in_dstep={k: v.grad.clone() for k, v in list(netg.named_parameters())[:6]},
            x_pos=x_pos.detach())
        # G step
        netg.zero_grad()
        for p in netd.parameters():
            p.requires_grad = False
        x0p = netg(x_tp1.detach(), t, z)
        mean = pc.posterior_mean_coef1[t].view(-1, 1, 1, 1) * x0p + pc.posterior_mean_coef2[t].view(-1, 1, 1, 1) * x_tp1
        x_pos = mean + (1 - (t == 0).float()).view(-1, 1, 1, 1) * torch.exp(0.5 * pc.posterior_log_variance_clipped[t].view(-1, 1, 1, 1)) * n_post
        errG = F.softplus(-netd(x_pos, t, x_tp1.detach()).view(-1)).mean()
        errG.backward()
        gG = {k: v.grad.clone() for k, v in netg.named_parameters()}
        keep = [k for k in gG if ('all_modules.3.' in k or 'all_modules.2.' in k or 'z_transform.1' in k
                                  or k.startswith('all_modules.0.') or 'all_modules.9.' in k)]
        out['train_tiny'].update(errG=errG.detach(), gradG={k: gG[k] for k in keep},
                                 gradG_norm=torch.sqrt(sum((v.double() ** 2).sum() for v in gG.values())))

    path = os.path.join(HERE, 'reference_golden.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path) / 1e6, 'MB')


if __name__ == '__main__':
    main()
