"""The reference-facing import surface (`score_sde.op.*`, `score_sde.models.*`): the names, signatures and autograd behaviour
train_ddgan.py / test_ddgan.py rely on, checked against the CPU oracle."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ddgan_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize('up,down,pad', [(2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (2, 2))])
def test_upfirdn2d_first_and_second_order_gradients(up, down, pad):
    """score_sde/op/upfirdn2d.py:27-164: differentiable to 2nd order w.r.t. input (the R1 penalty path), None grad for kernel."""
    from score_sde.op import upfirdn2d
    k = torch.from_numpy(O.setup_fir_kernel([1, 3, 3, 1])) * (up ** 2)
    x = seeded((2, 5, 12, 12), 1)
    gy_shape = tuple(O.upfirdn2d(x, k, up, down, pad).shape)
    gy = seeded(gy_shape, 2); v = seeded(tuple(x.shape), 3)

    def run(fn, xx, kk, g, vv):
        xx = xx.clone().requires_grad_(True)
        y = fn(xx, kk, up=up, down=down, pad=pad)
        gx, = torch.autograd.grad((y * g).sum() + (y ** 2).sum(), xx, create_graph=True)
        ggx, = torch.autograd.grad((gx * vv).sum(), xx)
        return y.detach(), gx.detach(), ggx

    y0, g0, gg0 = run(O.upfirdn2d, x, k, gy, v)
    y1, g1, gg1 = run(upfirdn2d, x.to(DEV), k.to(DEV), gy.to(DEV), v.to(DEV))
    assert O.rel_l2(y1.cpu(), y0) < 1e-5 and O.rel_l2(g1.cpu(), g0) < 1e-5 and O.rel_l2(gg1.cpu(), gg0) < 1e-5
    with pytest.raises(RuntimeError):
        upfirdn2d(x, k, up=up, down=down, pad=pad)          # CPU tensors: no fallback


def test_fused_leaky_relu_module_and_gradients():
    """score_sde/op/fused_act.py:82-105: FusedLeakyReLU(channel).bias is a Parameter[channel]; y = leaky_relu(x + b) * scale."""
    from score_sde.op import FusedLeakyReLU, fused_leaky_relu
    m = FusedLeakyReLU(6, negative_slope=0.2, scale=2 ** 0.5).to(DEV)
    assert isinstance(m.bias, torch.nn.Parameter) and tuple(m.bias.shape) == (6,)
    with torch.no_grad():
        m.bias.copy_(seeded((6,), 4, 0.3).to(DEV))
    x = seeded((3, 6, 7, 5), 5)
    xr = x.clone().requires_grad_(True); br = m.bias.detach().cpu().clone().requires_grad_(True)
    ref = F.leaky_relu(xr + br.view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    gy = seeded(tuple(ref.shape), 6)
    gx0, gb0 = torch.autograd.grad((ref * gy).sum(), (xr, br))
    xd = x.to(DEV).requires_grad_(True)
    y = m(xd)
    assert O.rel_l2(y.detach().cpu(), ref.detach()) < 1e-6
    gx1, gb1 = torch.autograd.grad((y * gy.to(DEV)).sum(), (xd, m.bias), create_graph=True)
    assert O.rel_l2(gx1.detach().cpu(), gx0) < 1e-6 and O.rel_l2(gb1.detach().cpu(), gb0) < 1e-5
    # double backward runs (the gate is piecewise constant: d/dx of sum(gx^2) is zero almost everywhere)
    ggy, = torch.autograd.grad((gx1 ** 2).sum(), xd, allow_unused=True)
    assert ggy is None or float(ggy.abs().max()) == 0.0
    y2 = fused_leaky_relu(x.to(DEV), m.bias.detach(), 0.1, 1.5)
    assert O.rel_l2(y2.cpu(), F.leaky_relu(x + br.detach().view(1, -1, 1, 1), 0.1) * 1.5) < 1e-6


def test_resampling_functions_match_reference_arithmetic():
    """score_sde/models/up_or_down_sampling.py:149-262 on the public NCHW surface."""
    from score_sde.models.up_or_down_sampling import conv_downsample_2d, downsample_2d, upsample_2d
    x = seeded((2, 32, 16, 16), 7)
    assert O.rel_l2(upsample_2d(x.to(DEV), (1, 3, 3, 1), factor=2).cpu(), O.upsample_2d(x)) < 1e-5
    assert O.rel_l2(downsample_2d(x.to(DEV), (1, 3, 3, 1), factor=2).cpu(), O.downsample_2d(x)) < 1e-5
    w = seeded((48, 32, 3, 3), 8) / math.sqrt(32 * 9)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    ref = O.conv_downsample_2d(xr, wr)
    gy = seeded(tuple(ref.shape), 9)
    g0 = torch.autograd.grad((ref * gy).sum(), (xr, wr))
    xd = x.to(DEV).requires_grad_(True); wd = w.to(DEV).requires_grad_(True)
    y = conv_downsample_2d(xd, wd, (1, 3, 3, 1), factor=2)
    assert tuple(y.shape) == tuple(ref.shape) and O.rel_l2(y.detach().cpu(), ref.detach()) < 2e-5
    g1 = torch.autograd.grad((y * gy.to(DEV)).sum(), (xd, wd))
    for a, b in zip(g1, g0):
        assert O.rel_l2(a.cpu(), b) < 5e-5


def test_model_classes_keep_reference_constructor_forward_and_state_dict(golden):
    """NCSNpp(config).forward(x, time_cond, z) and Discriminator_small(nc, ngf, t_emb_dim, act).forward(x, t, x_t) imported
    from the reference's module paths, loading reference-layout state dicts with strict=True (test_ddgan.py:162)."""
    from score_sde.models.discriminator import Discriminator_large, Discriminator_small
    from score_sde.models.ncsnpp_generator_adagn import NCSNpp
    cfg = O.tiny_config()
    netG = NCSNpp(cfg).to(DEV)
    sd = O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=7)
    assert list(netG.state_dict().keys()) == list(golden['ncsnpp_tiny_shapes'].keys())
    netG.load_state_dict(sd, strict=True)
    netG.eval()
    x = seeded((3, 3, 16, 16), 200); z = seeded((3, cfg.nz), 201); t = torch.tensor([0, 3, 1])
    with torch.no_grad():
        y = netG(x.to(DEV), t.to(DEV), z.to(DEV))
    assert O.rel_l2(y.cpu(), golden['ncsnpp_tiny_out']) < 1e-4
    # the same call under autograd takes the differentiable path and agrees with the fused one
    y2 = netG(x.to(DEV), t.to(DEV), z.to(DEV))
    assert y2.requires_grad and O.rel_l2(y2.detach().cpu(), y.cpu()) < 1e-4
    netD = Discriminator_small(nc=6, ngf=16, t_emb_dim=32, act=torch.nn.LeakyReLU(0.2)).to(DEV)
    sdd = O.randomize_params(O.discriminator_param_shapes(6, 16, 32), seed=22)
    netD.load_state_dict(sdd, strict=True)
    xa = seeded((4, 3, 32, 32), 210); xb = seeded((4, 3, 32, 32), 211); td = torch.tensor([0, 1, 2, 3])
    with torch.no_grad():
        d = netD(xa.to(DEV), td.to(DEV), xb.to(DEV))
    ref = O.discriminator_forward(sdd, xa, td, xb, 32)
    assert tuple(d.shape) == (4, 1) and O.rel_l2(d.cpu(), ref) < 1e-4
    assert Discriminator_large is not None


@pytest.mark.parametrize('dtype,tol', [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
def test_sixteen_bit_operator_surface(dtype, tol):
    """The reference dispatches both operators over fp16 as well (upfirdn2d_kernel.cu:313, fused_bias_act_kernel.cu:79); here fp16
    and bf16 tensors run 16-bit kernels with fp32 accumulation.  Expected values: the fp32 path on the upcast inputs; stated
    tolerance = rounding of the 16-bit output (bf16 2^-9, fp16 2^-11 relative), forward and first-order gradient."""
    from score_sde.op import upfirdn2d, fused_leaky_relu
    from score_sde.models.up_or_down_sampling import upsample_2d, downsample_2d
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 32, 16, 16, generator=g).to(DEV)
    xl = x.to(dtype)
    # 16 x 16 (width % 8 == 0): the register-tiled 16-bit fast paths; 20 x 12: the staged general kernel
    for xin in (xl, torch.randn(3, 8, 20, 12, generator=g).to(DEV).to(dtype), torch.randn(2, 4, 64, 64, generator=g).to(DEV).to(dtype)):
        for fn in (lambda t: upsample_2d(t, (1, 3, 3, 1), factor=2), lambda t: downsample_2d(t, (1, 3, 3, 1), factor=2)):
            ref = fn(xin.float())
            got = fn(xin)
            assert got.dtype == dtype and got.shape == ref.shape
            assert float((got.float() - ref).norm() / ref.norm()) < tol
    k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device=DEV) / 16
    xr = xl.clone().requires_grad_(True)
    xf = xl.float().requires_grad_(True)
    y = upfirdn2d(xr, k, up=1, down=1, pad=(1, 1))
    yf = upfirdn2d(xf, k, up=1, down=1, pad=(1, 1))
    gy = torch.randn(yf.shape, generator=g).to(DEV)
    y.backward(gy.to(dtype)); yf.backward(gy.to(dtype).float())
    assert float((xr.grad.float() - xf.grad).norm() / xf.grad.norm()) < tol
    b = torch.randn(32, generator=g).to(DEV)
    xr = xl.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    xf = xl.float().requires_grad_(True); bf = b.clone().requires_grad_(True)
    y = fused_leaky_relu(xr, br)
    yf = fused_leaky_relu(xf, bf)
    assert y.dtype == dtype and float((y.float() - yf).norm() / yf.norm()) < tol
    y.backward(gy.to(dtype)); yf.backward(gy.to(dtype).float())
    assert float((xr.grad.float() - xf.grad).norm() / xf.grad.norm()) < tol
    assert float((br.grad - bf.grad).norm() / bf.grad.norm()) < tol
