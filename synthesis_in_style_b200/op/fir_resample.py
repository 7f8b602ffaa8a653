"""upfirdn2d, host side.

Drop-in for scf/networks/stylegan2/op/upfirdn2d.py: `upfirdn2d(input[B,C,H,W], kernel[kh,kw], up=1, down=1, pad=(0,0))`
and the raw extension entry point `upfirdn2d_op(input[major,H,W,minor], kernel, up_x, up_y, down_x, down_y,
pad_x0, pad_x1, pad_y0, pad_y1)` (upfirdn2d.cpp:12-23), both running csrc/upfirdn2d.cu through the C-ABI.

Autograd.  For fixed taps the op is a linear map; its adjoint is the same op with the taps flipped, up and down swapped
and the pads mirrored (`_FirGeometry.adjoint`), and the adjoint of the adjoint is the original map.  ONE Function
(`_FirResample`) whose backward applies itself with the adjoint geometry therefore serves every derivative order.  (The
reference writes this as a Function pair with explicit gradient pads, upfirdn2d.py:18-84; the pad formulas agree.)
As in the reference the taps receive no gradient.
"""
from typing import NamedTuple, Tuple

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


class _FirGeometry(NamedTuple):
    """Resampling geometry of one application on [B, C, H, W] planes: (x, y) factors, (x0, x1, y0, y1) pads, plane sizes."""
    up: Tuple[int, int]
    down: Tuple[int, int]
    pad: Tuple[int, int, int, int]
    taps: Tuple[int, int]          # (kh, kw)
    src: Tuple[int, int]           # (H, W) of the input planes
    dst: Tuple[int, int]           # (H, W) of the output planes

    def adjoint(self) -> '_FirGeometry':
        """Geometry of the transposed map (dst-sized planes -> src-sized planes) for the flipped taps."""
        (ux, uy), (dx, dy), (px0, _, py0, _) = self.up, self.down, self.pad
        kh, kw = self.taps
        (h, w), (oh, ow) = self.src, self.dst
        pad = (kw - px0 - 1, w * ux - ow * dx + px0 - ux + 1, kh - py0 - 1, h * uy - oh * dy + py0 - uy + 1)
        return _FirGeometry(self.down, self.up, pad, self.taps, self.dst, self.src)


class _FirResample(Function):
    @staticmethod
    def forward(ctx, planes, taps, geo: _FirGeometry):
        lead = planes.shape[:-2]
        out = upfirdn2d_op(planes.reshape(-1, geo.src[0], geo.src[1], 1), taps, geo.up[0], geo.up[1], geo.down[0], geo.down[1],
                           *geo.pad)
        assert tuple(out.shape[1:3]) == tuple(geo.dst), (out.shape, geo)
        ctx.save_for_backward(taps)
        ctx.geo = geo
        return out.reshape(*lead, geo.dst[0], geo.dst[1])

    @staticmethod
    def backward(ctx, g):
        taps, = ctx.saved_tensors
        return _FirResample.apply(g, torch.flip(taps, [0, 1]), ctx.geo.adjoint()), None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    lib = _lib.load()
    kh, kw = kernel.shape
    h, w = input.shape[-2:]
    dst = (lib.sis_upfirdn2d_out_size(h, up, down, pad[0], pad[1], kh), lib.sis_upfirdn2d_out_size(w, up, down, pad[0], pad[1], kw))
    geo = _FirGeometry((up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]), (kh, kw), (h, w), dst)
    return _FirResample.apply(input, kernel, geo)
