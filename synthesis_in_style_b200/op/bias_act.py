"""fused bias + leaky ReLU, host side.

Drop-in for scf/networks/stylegan2/op/fused_act.py: `fused_leaky_relu(input, bias, negative_slope=0.2, scale=2**0.5)`,
`FusedLeakyReLU(channel, negative_slope=0.2, scale=2**0.5)` and the raw extension entry point
`fused_bias_act(input, bias, refer, act, grad, alpha, scale)` (fused_bias_act.cpp:11-20), all running the sm_100a
kernel in csrc/fused_bias_act.cu through the C-ABI.

Autograd.  y = lrelu(x + b) * scale.  Its Jacobian w.r.t. x is diagonal: g -> g * gate(y) * scale with gate = 1 where
y > 0 and `negative_slope` elsewhere -- the kernel's (act=3, grad=1) mode with `ref = y`.  A diagonal map is its own
adjoint, so ONE Function (`_LeakyGate`) whose backward applies itself again serves every derivative order; the bias
gradient is a plain `sum` that autograd differentiates by itself.  (The reference spells the same mathematics as a
Function pair with a hand-written double-backward, fused_act.py:19-70.)
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import _lib

_DTYPES = {torch.float32: _lib.SIS_F32, torch.float16: _lib.SIS_F16, torch.float64: _lib.SIS_F64}
_ACT_LRELU, _MODE_VALUE, _MODE_GATE = 3, 0, 1


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


class _LeakyGate(Function):
    """g -> g * gate(y) * scale.  Linear in g and self-adjoint, so `backward` is the Function itself."""

    @staticmethod
    def forward(ctx, g, y, slope, scale):
        ctx.save_for_backward(y)
        ctx.slope, ctx.scale = slope, scale
        return fused_bias_act(g, g.new_empty(0), y, _ACT_LRELU, _MODE_GATE, slope, scale)

    @staticmethod
    def backward(ctx, gg):
        y, = ctx.saved_tensors
        return _LeakyGate.apply(gg, y, ctx.slope, ctx.scale), None, None, None


class _BiasLeakyReLU(Function):
    @staticmethod
    def forward(ctx, x, bias, slope, scale):
        y = fused_bias_act(x, bias, x.new_empty(0), _ACT_LRELU, _MODE_VALUE, slope, scale)
        ctx.save_for_backward(y)
        ctx.slope, ctx.scale, ctx.has_bias = slope, scale, bias.numel() > 0
        return y

    @staticmethod
    def backward(ctx, g):
        y, = ctx.saved_tensors
        gx = _LeakyGate.apply(g, y, ctx.slope, ctx.scale)
        gb = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            gb = gx.sum([d for d in range(gx.ndim) if d != 1])       # bias broadcasts over every dim but 1
        return gx, gb, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return _BiasLeakyReLU.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
