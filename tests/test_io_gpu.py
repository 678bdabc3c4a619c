"""Image-output and checkpoint formats (test_ddgan.py:190-201, ddgan.py:373-410, 545-569)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import ddgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def test_images_to_uint8_matches_save_image_quantisation(tmp_path):
    from ddgan_b200 import io
    x = (seeded((5, 3, 32, 32), 1) * 0.8).clamp(-1.3, 1.3)
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.999])
    u8 = io.images_to_uint8(x.to(DEV)).cpu()
    # test_ddgan.py:149 to_range_0_1, then torchvision.utils.save_image: mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(uint8)
    ref = ((x + 1.0) / 2.0).mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8)
    assert (u8.int() - ref.int()).abs().max() <= 1          # fp32 rounding of (x+1)/2*255 vs x*0.5+0.5 may differ by one level
    assert (u8 != ref).float().mean() < 1e-3
    w = io.ImageWriter(str(tmp_path), fmt='png', workers=4, save_npy=True)
    w.write(x.to(DEV), start_index=10)
    w.close()
    from PIL import Image
    for j in (0, 4):
        img = np.array(Image.open(os.path.join(str(tmp_path), f'{10 + j}.png')))
        assert img.shape == (32, 32, 3) and np.array_equal(img, u8[j].numpy())
        f = np.load(os.path.join(str(tmp_path), f'{10 + j}.npy'))
        assert np.allclose(f, ((x[j] + 1) / 2).numpy(), atol=1e-6)


def _nets():
    from ddgan_b200.modules import NCSNpp, Discriminator_small
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
    netG = NCSNpp(cfg).to(DEV)
    netG.load_state_dict(O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21))
    netD = Discriminator_small(nc=6, ngf=16, t_emb_dim=32).to(DEV)
    netD.load_state_dict(O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22))
    for k, v in dict(lr_g=1.6e-4, lr_d=0.0, beta1_g=0.5, beta2_g=0.9, beta1_d=0.5, beta2_d=0.9, r1_gamma=0.02, lazy_reg=2,
                     grad_clip_norm=1.0, ema_decay=0.999, use_ema=True).items():
        setattr(cfg, k, v)
    return cfg, netG, netD


def _noise(cfg, B, base):
    nz = {}
    for sfx, b in (('_d', base), ('_g', base + 50)):
        nz['t' + sfx] = (torch.arange(B) % cfg.num_timesteps).to(DEV)
        for j, k in enumerate(('n_xtp1', 'n_xt', 'n_post')):
            nz[k + sfx] = seeded((B, 3, 32, 32), b + 1 + j).to(DEV)
        nz['z' + sfx] = seeded((B, cfg.nz), b + 9).to(DEV)
    return nz


def test_checkpoint_round_trip_in_reference_layout(tmp_path):
    """Train 2 steps, write content.pth + netG_<epoch>.pth, resume in a fresh Trainer, train 1 more step: weights, Adam state and
    EMA continue exactly as in the uninterrupted run (lr_d = 0 keeps the G step deterministic, see
    test_parity_baseline_gpu.py); the file has the reference's keys and its generator dict loads strictly into a fresh module."""
    from ddgan_b200 import io
    from ddgan_b200.modules import NCSNpp
    from ddgan_b200.train import Trainer
    cfg, netG, netD = _nets()
    B = 4
    real = torch.tanh(seeded((B, 3, 32, 32), 300)).to(DEV)
    a = Trainer(cfg, netG, netD, DEV)
    for it in range(2):
        a.step(real, it, noise=_noise(cfg, B, 800 + 10 * it))
    path = os.path.join(str(tmp_path), 'content.pth')
    content = io.save_checkpoint(path, a, epoch=3, global_step=2)
    assert set(content) >= {'epoch', 'global_step', 'args', 'netG_dict', 'optimizerG', 'netD_dict', 'optimizerD', 'emaG'}
    assert all(v.device.type == 'cpu' for v in content['emaG'].values())
    io.save_generator(os.path.join(str(tmp_path), 'netG_3.pth'), a)
    # resume
    cfg2, netG2, netD2 = _nets()
    for p in list(netG2.parameters()) + list(netD2.parameters()):
        p.data.normal_()                                   # whatever was there is overwritten by the checkpoint
    b = Trainer(cfg2, netG2, netD2, DEV)
    epoch, gs, _ = io.load_checkpoint(path, b, map_location=DEV)
    assert (epoch, gs) == (3, 2)
    assert torch.equal(b.optG.flat_p, a.optG.flat_p) and torch.equal(b.optG.m, a.optG.m) and torch.equal(b.optG.ema, a.optG.ema)
    assert float(b.optG.state[0]) == 2.0
    nz = _noise(cfg, B, 820)
    ea = a.step(real, 2, noise=nz)
    eb = b.step(real, 2, noise=nz)
    assert abs(float(ea[1]) - float(eb[1])) < 1e-5 * max(1.0, abs(float(ea[1])))
    assert O.rel_l2(b.optG.flat_p.cpu(), a.optG.flat_p.cpu()) < 1e-5
    assert O.rel_l2(b.optG.ema.cpu(), a.optG.ema.cpu()) < 1e-5
    # netG_<epoch>.pth = EMA weights, strict-loadable, and the live parameters were swapped back
    sd = torch.load(os.path.join(str(tmp_path), 'netG_3.pth'), map_location=DEV)
    fresh = NCSNpp(cfg).to(DEV)
    fresh.load_state_dict(sd, strict=True)
    ema = content['emaG']
    k0 = next(iter(ema))
    assert torch.allclose(sd[k0].cpu(), ema[k0])
    # the reference's torch.optim.Adam accepts the optimiser entry as it is
    torch.optim.Adam(copy.deepcopy(fresh).parameters(), lr=1e-4).load_state_dict(content['optimizerG'])
