"""1-vs-N GPU equivalence on hardware (ddgan.py:363-365 DistributedDataParallel): two NCCL ranks, each with
its own half of a global batch, must end the step with the gradient a single process gets by running both halves and summing
(the flat arena is all-reduced with SUM; the mean's 1/world rides in the optimiser pass), and with identical parameters on
every rank after a real update.  Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, 'denoising-diffusion-gan_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import copy
    import torch.distributed as dist
    from oracle import ddgan_oracle as O
    from ddgan_b200.modules import NCSNpp, Discriminator_small
    from ddgan_b200.train import Trainer
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)

    def seeded(shape, seed):
        return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
    cfg = O.tiny_config(image_size=32, attn_resolutions=(16,), t_emb_dim=32, ngf=16)
    for k, v in dict(lr_g=0.0, lr_d=0.0, beta1_g=0.5, beta2_g=0.9, beta1_d=0.5, beta2_d=0.9, r1_gamma=0.02, lazy_reg=1,
                     grad_clip_norm=1.0, ema_decay=0.999, use_ema=True).items():
        setattr(cfg, k, v)
    torch.manual_seed(100 + rank)                        # different init per rank: the Trainer's broadcast must equalise
    netG = NCSNpp(cfg).to(dev); netD = Discriminator_small(nc=6, ngf=16, t_emb_dim=32).to(dev)
    if rank == 0:
        netG.load_state_dict(O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=21))
        netD.load_state_dict(O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22))
    tr = Trainer(cfg, netG, netD, dev, distributed=True)
    netG1, netD1 = copy.deepcopy(netG), copy.deepcopy(netD)      # after the broadcast: rank 0's weights everywhere
    B = 4                                                # per rank

    def noise(r, base):
        nz = {}
        for sfx, b in (('_d', base + 100 * r), ('_g', base + 100 * r + 50)):
            nz['t' + sfx] = (torch.arange(B) % cfg.num_timesteps).to(dev)
            for j, k in enumerate(('n_xtp1', 'n_xt', 'n_post')):
                nz[k + sfx] = seeded((B, 3, 32, 32), b + 1 + j).to(dev)
            nz['z' + sfx] = seeded((B, cfg.nz), b + 9).to(dev)
        return nz
    reals = [torch.tanh(seeded((B, 3, 32, 32), 300 + r)).to(dev) for r in range(world)]
    tr.step(reals[rank], 0, noise=noise(rank, 1000))
    def by_name(opt):
        # arenas of a distributed and a single-process trainer are laid out differently (early / late ranges): compare by name
        return {n: opt.flat_g[off:off + num].clone() for n, (off, num) in zip(opt.names, opt.views)}
    gD, gG = by_name(tr.optD), by_name(tr.optG)
    assert tr.optG.n_early > 0
    # single-process reference: both halves one after the other on a non-distributed trainer with the same weights (lr = 0)
    ref = Trainer(cfg, netG1, netD1, dev, distributed=False)
    sD = {k: torch.zeros_like(v) for k, v in gD.items()}
    sG = {k: torch.zeros_like(v) for k, v in gG.items()}
    for r in range(world):
        ref.step(reals[r], 0, noise=noise(r, 1000))
        for k, v in by_name(ref.optD).items():
            sD[k] += v
        for k, v in by_name(ref.optG).items():
            sG[k] += v
    cat = lambda d: torch.cat([d[k].flatten() for k in sorted(d)])
    errD, errG = O.rel_l2(cat(gD).cpu(), cat(sD).cpu()), O.rel_l2(cat(gG).cpu(), cat(sG).cpu())
    # a real update: every rank must hold identical parameters afterwards
    tr.optD.set_lr(1.25e-4); tr.optG.set_lr(1.6e-4)
    tr.step(reals[rank], 1, noise=noise(rank, 2000))
    chk = torch.stack([tr.optD.flat_p.double().sum(), tr.optG.flat_p.double().sum(), tr.optG.flat_p.double().abs().sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(bool(torch.equal(allc[0], c)) for c in allc)
    q.put((rank, errD, errG, same, float(tr.optG.grad_scale)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_on_both_halves():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, errD, errG, same, scale in res:
        assert errD < 1e-5 and errG < 1e-5, (rank, errD, errG)   # summation order of the all-reduce / atomics only
        assert same and scale == 0.5
