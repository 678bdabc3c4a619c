"""score_sde.op.fused_leaky_relu / FusedLeakyReLU drop-in (reference: score_sde/op/fused_act.py:28-105).

y = leaky_relu(x + bias[c], negative_slope) * scale through ddg_fused_bias_act (act=3); the backward gates by the saved
output (grad=1) and reduces grad_bias with ddg_channel_sum; the double backward reuses the gated kernel, as in the
reference.  `negative_slope` is honoured (the reference's CPU branch hard-codes 0.2, fused_act.py:99)."""
import torch
from torch import nn
from torch.autograd import Function

from ddgan_b200 import ops


class FusedLeakyReLUFunctionBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = ops.fused_bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
        gsum = grad_input if grad_input.dtype == torch.float32 else grad_input.float()   # 16-bit gradients: reduce in fp32
        grad_bias = ops.channel_sum(gsum) if gsum.ndim >= 2 else gsum.sum(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = ops.fused_bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = ops.fused_bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output.contiguous(), out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    if input.device.type != 'cuda':
        raise RuntimeError('score_sde.op.fused_leaky_relu (ddgan_b200): CUDA tensors only, there is no CPU fallback')
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
