"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: the flat gradient all-reduce reproduces the
full-batch gradient (the rule SURVEY.md 2.4 measured for the reference's DDP), and sampling shards by rank with seed+rank."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ddgan_b200.train import FlatGradAllReducer, broadcast_params
    torch.manual_seed(1234 + rank)                      # different init per rank -> broadcast must equalise
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))
    broadcast_params(net.parameters())
    g = torch.Generator().manual_seed(7)
    data = torch.randn(8, 8, generator=g)               # global batch, identical on both ranks
    local = data[rank * 4:(rank + 1) * 4]
    ar = FlatGradAllReducer(net.parameters())
    # three accumulating backward calls, like the D step (real, R1, fake)
    net(local).mean().backward()
    (net(local) ** 2).mean().backward()
    net(local * 2).mean().backward()
    ar.allreduce()
    grads = [p.grad.clone() for p in net.parameters()]
    # full-batch reference on every rank (loss means are over the local batch, so the global mean = mean of rank means)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))
    ref.load_state_dict(net.state_dict())
    (ref(data).mean() + (ref(data) ** 2).mean() + ref(data * 2).mean()).backward()
    err = max(float((a - b.grad).abs().max()) for a, b in zip(grads, ref.parameters()))
    # The fused path (train.FlatAdam, what bench.py times): parameters and .grad are views into flat arenas, the collective is
    # ONE sum all-reduce of the gradient arena, and the mean's 1/world rides in the optimiser pass as `grad_scale`.
    from ddgan_b200.train import FlatAdam
    net2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))
    net2.load_state_dict(ref.state_dict())
    fa = FlatAdam(net2, 1e-3, (0.5, 0.999), max_norm=1.0)
    fa.grad_scale = 1.0 / world
    fa.zero_grad()
    net2(local).mean().backward()
    (net2(local) ** 2).mean().backward()
    net2(local * 2).mean().backward()
    assert all(p.grad.data_ptr() == fa.flat_g.data_ptr() + 4 * off for p, (off, _) in zip(fa.params, fa.views))
    dist.all_reduce(fa.flat_g)
    err2 = max(float((p.grad * fa.grad_scale - b.grad).abs().max()) for p, b in zip(net2.parameters(), ref.parameters()))
    q.put((rank, max(err, err2), float(sum(p.sum() for p in net.parameters()))))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_matches_full_batch_gradient():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] < 1e-6 and res[1][1] < 1e-6
    assert abs(res[0][2] - res[1][2]) < 1e-6            # parameters identical after broadcast
