"""One small invocation of the hot path on cuda:0, checked against the CPU oracle (used by __graft_entry__.smoke())."""
import os
import sys

import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import ddgan_oracle as O  # checker only
    from .engine import GeneratorEngine
    from . import diffusion
    cfg = O.tiny_config()
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=3)
    B = 2
    g = torch.Generator().manual_seed(4)
    x_init = torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=g)
    draws = []

    def noise_fn(shape):
        d = torch.randn(*shape, generator=g)
        draws.append(d)
        return d
    ref = O.sample_from_model(O.posterior_coefficients(cfg), lambda x, t, z: O.ncsnpp_forward(sd, cfg, x, t, z),
                              cfg.num_timesteps, x_init, cfg.nz, noise_fn)
    eng = GeneratorEngine(cfg, B, 'cuda:0')
    eng.load_state_dict(sd)
    it = iter(draws)
    y = diffusion.sample_from_model(diffusion.PosteriorCoefficients(cfg, 'cuda:0'), eng.forward, cfg.num_timesteps,
                                    x_init.cuda(), None, cfg, noise_fn=lambda s: next(it).cuda())
    err = O.rel_l2(y.cpu(), ref)
    print(f'smoke: T={cfg.num_timesteps} sampling of {B} images, rel-L2 vs oracle = {err:.3e}, '
          f'{eng.n_launches} launches per generator forward')
    assert err < 1e-4, err
