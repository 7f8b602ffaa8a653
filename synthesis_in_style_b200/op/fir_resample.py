"""upfirdn2d, host side.

Mirrors scf/networks/stylegan2/op/upfirdn2d.py: `upfirdn2d(input[B,C,H,W], kernel[kh,kw], up=1, down=1, pad=(0,0))`
and the raw extension entry point `upfirdn2d_op(input[major,H,W,minor], kernel, up_x, up_y, down_x, down_y,
pad_x0, pad_x1, pad_y0, pad_y1)` (upfirdn2d.cpp:12-23), both running csrc/upfirdn2d.cu through the C-ABI.
The gradient is the same op with the flipped kernel and swapped up/down (upfirdn2d.py:18-84).
"""
import torch
from torch.autograd import Function

from .. import _lib

_DTYPES = {torch.float32: _lib.SIS_F32, torch.float16: _lib.SIS_F16, torch.float64: _lib.SIS_F64}


def upfirdn2d_op(input: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int, pad_x0: int,
                 pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    _lib.require_cuda(input, 'input')
    _lib.require_cuda(kernel, 'kernel')
    if input.dtype not in _DTYPES:
        raise RuntimeError(f'"upfirdn2d_cuda" not implemented for \'{input.dtype}\'')
    if input.dim() != 4 or kernel.dim() != 2:
        raise RuntimeError('upfirdn2d: input must be [major, H, W, minor] and kernel [kh, kw]')
    x = input.contiguous()
    k = kernel.contiguous().to(x.dtype)
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    lib = _lib.load()
    out_h = lib.sis_upfirdn2d_out_size(in_h, up_y, down_y, pad_y0, pad_y1, kh)
    out_w = lib.sis_upfirdn2d_out_size(in_w, up_x, down_x, pad_x0, pad_x1, kw)
    if out_h < 0 or out_w < 0:
        raise RuntimeError(f'upfirdn2d: negative output size {out_h}x{out_w}')
    out = torch.empty((major, out_h, out_w, minor), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.sis_upfirdn2d(_lib.ptr(out), _lib.ptr(x), _lib.ptr(k), _DTYPES[x.dtype], major, in_h, in_w, minor,
                                     kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1,
                                     _lib.current_stream_ptr(x.device)))
    return out


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        up_x, up_y = up
        down_x, down_y = down
        g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1 = g_pad
        grad_output = grad_output.reshape(-1, out_size[0], out_size[1], 1)
        grad_input = upfirdn2d_op(grad_output, grad_kernel, down_x, down_y, up_x, up_y, g_pad_x0, g_pad_x1, g_pad_y0,
                                  g_pad_y1)
        grad_input = grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gradgrad_input = gradgrad_input.reshape(-1, ctx.in_size[2], ctx.in_size[3], 1)
        out = upfirdn2d_op(gradgrad_input, kernel, ctx.up[0], ctx.up[1], ctx.down[0], ctx.down[1], *ctx.pad)
        out = out.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1])
        return out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        batch, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        out = upfirdn2d_op(input.reshape(-1, in_h, in_w, 1), kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0,
                           pad_y1)
        out_h, out_w = out.shape[1], out.shape[2]
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        ctx.g_pad = (kernel_w - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kernel_h - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad,
                                             ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
