"""fused bias + leaky ReLU, host side.

Mirrors scf/networks/stylegan2/op/fused_act.py: `fused_leaky_relu(input, bias, negative_slope=0.2, scale=2**0.5)`,
`FusedLeakyReLU(channel, negative_slope=0.2, scale=2**0.5)` and the raw extension entry point
`fused_bias_act(input, bias, refer, act, grad, alpha, scale)` (fused_bias_act.cpp:11-20), all running the sm_100a
kernel in csrc/fused_bias_act.cu through the C-ABI.  Autograd (first and second order) follows the reference's
two Function classes with the same kernel modes (act=3, grad=0/1).
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import _lib

_DTYPES = {torch.float32: _lib.SIS_F32, torch.float16: _lib.SIS_F16, torch.float64: _lib.SIS_F64}


def fused_bias_act(input: torch.Tensor, bias: torch.Tensor, refer: torch.Tensor, act: int, grad: int, alpha: float,
                   scale: float) -> torch.Tensor:
    """`fused.fused_bias_act` (fused_bias_act.cpp:11-20): returns a new tensor, launches on the current stream."""
    _lib.require_cuda(input, 'input')
    _lib.require_cuda(bias, 'bias')
    if input.dtype not in _DTYPES:
        raise RuntimeError(f'"fused_bias_act_kernel" not implemented for \'{input.dtype}\'')
    x = input.contiguous()
    b = bias.contiguous().to(x.dtype)
    ref = refer.contiguous().to(x.dtype) if refer is not None and refer.numel() else None
    step_b = 1
    for i in range(2, x.dim()):
        step_b *= x.size(i)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().sis_fused_bias_act(
            _lib.ptr(y), _lib.ptr(x), _lib.ptr(b if b.numel() else None), _lib.ptr(ref), _DTYPES[x.dtype], x.numel(),
            step_b, b.numel(), int(act), int(grad), float(alpha), float(scale), _lib.current_stream_ptr(x.device)))
    return y


class FusedLeakyReLUFunctionBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        empty = grad_output.new_empty(0)
        grad_input = fused_bias_act(grad_output, empty, out, 3, 1, negative_slope, scale)
        dims = [0] + list(range(2, grad_input.ndim))
        grad_bias = grad_input.sum(dims).detach()
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = fused_bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = fused_bias_act(input, bias, input.new_empty(0), 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
