"""score_sde.op.upfirdn2d drop-in (reference: score_sde/op/upfirdn2d.py:27-164).

Same public signature `upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`, differentiable to any order w.r.t. `input`:
the op family is closed under differentiation -- the adjoint of (up, down, pad, k) is
(down, up, g_pad, flip(k)) with g_pad from upfirdn2d.py:119-122, and the adjoint of the adjoint is the forward op -- so
one kernel entry (ddg_upfirdn2d) serves forward, backward and double-backward (needed by the lazy R1 penalty).
There is no CPU fallback: CPU tensors raise."""
import torch
from torch.autograd import Function

from ddgan_b200 import ops


def _apply(x4, kernel, up, down, pad):
    n, c, h, w = x4.shape
    out = ops.upfirdn2d_raw(x4.reshape(n * c, h, w), kernel, up[0], up[1], down[0], down[1], pad[0], pad[1], pad[2], pad[3])
    return out.view(n, c, out.shape[1], out.shape[2])


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad, ctx.in_size, ctx.out_size = up, down, pad, in_size, out_size
        go = grad_output.reshape(in_size[0], in_size[1], out_size[0], out_size[1])
        gi = _apply(go, grad_kernel, down, up, g_pad)
        return gi.view(in_size)

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        ggo = UpFirDn2d.apply(gradgrad_input.reshape(ctx.in_size), kernel, ctx.up, ctx.down, ctx.pad)
        return ggo, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        _, _, in_h, in_w = input.shape
        ctx.in_size = input.shape
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) // down_y + 1
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) // down_x + 1
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        # upfirdn2d.py:119-122
        ctx.g_pad = (kernel_w - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kernel_h - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        return _apply(input, kernel, ctx.up, ctx.down, ctx.pad)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        gi = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad, ctx.in_size,
                                     ctx.out_size)
        return gi, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if input.device.type != 'cuda':
        raise RuntimeError('score_sde.op.upfirdn2d (ddgan_b200): CUDA tensors only, there is no CPU fallback')
    kernel = kernel.to(device=input.device, dtype=torch.float32)
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
