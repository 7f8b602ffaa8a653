"""CPU oracle: StyleGAN2 generator forward with activation capture.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch-CPU fp32 restatement
of `scf/networks/stylegan2/model.py` and of the two CUDA ops in
`scf/networks/stylegan2/op/`.  The generator is expressed functionally over a
state dict with the reference's key names, so that weights are
interchangeable with the product's `Generator` and with `g_ema` checkpoints.

Pinned by `tests/golden/make_golden.py` against the reference's own `model.py`
run in the build container (bit-for-bit on CPU, same ATen ops in the same
order).
"""
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- ops

def fused_bias_act(x: torch.Tensor, b: torch.Tensor, ref: torch.Tensor, act: int, grad: int,
                   alpha: float, scale: float) -> torch.Tensor:
    """`fused_bias_act_kernel`, scf/networks/stylegan2/op/fused_bias_act_kernel.cu:18-49.

    y = act(x + b[(i / step_b) % size_b]) * scale with step_b = prod(dims[2:]) (:67-71);
    act*10+grad: 10/11 linear, 12 zero, 30 lrelu, 31 lrelu gated by sign of `ref`, 32 zero.
    Anything else falls to the `default:` label = linear.
    """
    x = x.contiguous()
    if b.numel():
        shape = [1, -1] + [1] * (x.dim() - 2)
        x = x + b.reshape(shape).to(x.dtype)
    code = act * 10 + grad
    a = torch.tensor(alpha, dtype=x.dtype)
    if code == 30:
        y = torch.where(x > 0, x, x * a)
    elif code == 31:
        y = torch.where(ref > 0, x, x * a)
    elif code in (12, 32):
        y = torch.zeros_like(x)
    else:
        y = x
    return y * torch.tensor(scale, dtype=x.dtype)


def fused_leaky_relu(x: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2,
                     scale: float = 2 ** 0.5) -> torch.Tensor:
    """`fused_leaky_relu`, scf/networks/stylegan2/op/fused_act.py:85-86 (act=3, grad=0)."""
    return fused_bias_act(x, bias, x.new_empty(0), 3, 0, negative_slope, scale)


def upfirdn2d_op(x: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
                 pad_x0: int, pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    """`upfirdn2d_native`, scf/networks/stylegan2/op/upfirdn2d.py:152-186, on [major, H, W, minor].

    Zero-insert upsample, pad (negative pad crops), true convolution with the
    flipped taps, decimate.  Same semantics as the CUDA kernel
    (upfirdn2d_kernel.cu:71-133).
    """
    _, in_h, in_w, minor = x.shape
    kernel_h, kernel_w = kernel.shape
    out = x.reshape(-1, in_h, 1, in_w, 1, minor)
    out = F.pad(out, [0, 0, 0, up_x - 1, 0, 0, 0, up_y - 1])
    out = out.reshape(-1, in_h * up_y, in_w * up_x, minor)
    out = F.pad(out, [0, 0, max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    out = out[:, max(-pad_y0, 0): out.shape[1] - max(-pad_y1, 0), max(-pad_x0, 0): out.shape[2] - max(-pad_x1, 0), :]
    out = out.permute(0, 3, 1, 2)
    out = out.reshape([-1, 1, in_h * up_y + pad_y0 + pad_y1, in_w * up_x + pad_x0 + pad_x1])
    w = torch.flip(kernel, [0, 1]).view(1, 1, kernel_h, kernel_w)
    out = F.conv2d(out, w)
    out = out.reshape(-1, minor, in_h * up_y + pad_y0 + pad_y1 - kernel_h + 1,
                      in_w * up_x + pad_x0 + pad_x1 - kernel_w + 1)
    out = out.permute(0, 2, 3, 1)
    return out[:, ::down_y, ::down_x, :]


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up: int = 1, down: int = 1, pad=(0, 0)) -> torch.Tensor:
    """`upfirdn2d` / `UpFirDn2d.forward`, scf/networks/stylegan2/op/upfirdn2d.py:87-149 on NCHW."""
    b, c, h, w = x.shape
    out = upfirdn2d_op(x.reshape(-1, h, w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return out.reshape(-1, c, out.shape[1], out.shape[2])


def upfirdn2d_index_emulation(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """Pure-Python restatement of the CUDA kernel's index math
    (upfirdn2d_kernel.cu:71-133: tap flip, floor_div, kernel phase), small inputs only.
    Used to show the `upfirdn2d_native` restatement and the kernel agree."""
    import numpy as np
    x = np.asarray(x, dtype=np.float64)
    k = np.asarray(kernel, dtype=np.float64)
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    sk = k[::-1, ::-1]
    out = np.zeros((major, out_h, out_w, minor))
    for oy in range(out_h):
        mid_y = oy * down_y + up_y - 1 - pad_y0
        in_y0 = mid_y // up_y  # python // is floor_div
        ky0 = (in_y0 + 1) * up_y - mid_y - 1
        for ox in range(out_w):
            mid_x = ox * down_x + up_x - 1 - pad_x0
            in_x0 = mid_x // up_x
            kx0 = (in_x0 + 1) * up_x - mid_x - 1
            acc = np.zeros((major, minor))
            y = 0
            while ky0 + y * up_y < kh:
                xx = 0
                while kx0 + xx * up_x < kw:
                    iy, ix = in_y0 + y, in_x0 + xx
                    if 0 <= iy < in_h and 0 <= ix < in_w:
                        acc += x[:, iy, ix, :] * sk[ky0 + y * up_y, kx0 + xx * up_x]
                    xx += 1
                y += 1
            out[:, oy, ox, :] = acc
    return out


def make_kernel(k: Sequence[float]) -> torch.Tensor:
    """`make_kernel`, scf/networks/stylegan2/model.py:23-31."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


# ------------------------------------------------------------------ architecture

def get_channels(channel_multiplier: int = 2) -> Dict[int, int]:
    """`Generator.get_channels`, model.py:443-455."""
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier, 128: 128 * channel_multiplier,
            256: 64 * channel_multiplier, 512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}


class GeneratorSpec:
    """Static shape bookkeeping of `Generator.__init__`, model.py:367-441."""

    def __init__(self, size: int, style_dim: int, n_mlp: int, channel_multiplier: int = 2,
                 blur_kernel=(1, 3, 3, 1), lr_mlp: float = 0.01):
        self.size, self.style_dim, self.n_mlp = size, style_dim, n_mlp
        self.channel_multiplier = channel_multiplier
        self.blur_kernel = list(blur_kernel)
        self.lr_mlp = lr_mlp
        self.channels = get_channels(channel_multiplier)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.n_latent = self.log_size * 2 - 2


def init_state_dict(spec: GeneratorSpec, seed: Optional[int] = None) -> "OrderedDict[str, torch.Tensor]":
    """Random init drawing from the CPU generator in the SAME ORDER as the reference's
    constructors (model.py:139 EqualLinear, :223-227 ModulatedConv2d, :299 ConstantInput,
    :285 NoiseInjection zeros, :353 ToRGB bias zeros, :410-413 noise buffers before the conv loop),
    so `torch.manual_seed(s)` gives the reference's weights exactly."""
    if seed is not None:
        torch.manual_seed(seed)
    sd = OrderedDict()
    sdim = spec.style_dim
    for i in range(spec.n_mlp):
        sd[f'style.{i + 1}.weight'] = torch.randn(sdim, sdim).div_(spec.lr_mlp)
        sd[f'style.{i + 1}.bias'] = torch.zeros(sdim)
    c4 = spec.channels[4]
    sd['input.input'] = torch.randn(1, c4, 4, 4)

    def modconv(prefix, cin, cout, k, with_blur):
        sd[f'{prefix}.weight'] = torch.randn(1, cout, cin, k, k)
        if with_blur:
            sd[f'{prefix}.blur.kernel'] = make_kernel(spec.blur_kernel) * 4
        sd[f'{prefix}.modulation.weight'] = torch.randn(cin, sdim)
        sd[f'{prefix}.modulation.bias'] = torch.ones(cin)

    def styled(prefix, cin, cout, up):
        modconv(f'{prefix}.conv', cin, cout, 3, up)
        sd[f'{prefix}.noise.weight'] = torch.zeros(1)
        sd[f'{prefix}.activate.bias'] = torch.zeros(cout)

    def torgb(prefix, cin, up):
        if up:
            sd[f'{prefix}.upsample.kernel'] = make_kernel(spec.blur_kernel) * 4
        modconv(f'{prefix}.conv', cin, 3, 1, False)
        sd[f'{prefix}.bias'] = torch.zeros(1, 3, 1, 1)

    styled('conv1', c4, c4, False)
    torgb('to_rgb1', c4, False)
    for layer_idx in range(spec.num_layers):
        res = (layer_idx + 5) // 2
        sd[f'noises.noise_{layer_idx}'] = torch.randn(1, 1, 2 ** res, 2 ** res)
    cin = c4
    for j, i in enumerate(range(3, spec.log_size + 1)):
        cout = spec.channels[2 ** i]
        styled(f'convs.{2 * j}', cin, cout, True)
        styled(f'convs.{2 * j + 1}', cout, cout, False)
        torgb(f'to_rgbs.{j}', cout, True)
        cin = cout
    return sd


def reorder_like_module(sd):
    """nn.Module.state_dict() lists parameters before buffers per module; key ORDER is irrelevant
    for the oracle (lookups are by name)."""
    return sd


def perturb_zero_params(sd, seed: int = 1234, std: float = 0.1):
    """default init has noise weights / biases == 0, which hides the noise and bias
    paths.  For parity runs perturb them (own torch.Generator, does not touch the global RNG)."""
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if k.endswith('noise.weight') or k.endswith('activate.bias') or (k.endswith('.bias') and sd[k].dim() == 4) \
                or (k.startswith('style.') and k.endswith('.bias')):
            sd[k] = torch.randn(sd[k].shape, generator=g) * std
    return sd


# --------------------------------------------------------------------- forward

def pixel_norm(x):
    """`PixelNorm.forward`, model.py:19-20."""
    return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """`EqualLinear.forward`, model.py:152-162; scale = lr_mul / sqrt(in_dim) (:149)."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        out = F.linear(x, weight * scale)
        return fused_leaky_relu(out, bias * lr_mul)
    return F.linear(x, weight * scale, bias=bias * lr_mul)


def style_mlp(sd, spec: GeneratorSpec, z):
    """`Generator.style`, model.py:383-392 (PixelNorm + n_mlp EqualLinear with fused lrelu)."""
    x = pixel_norm(z)
    for i in range(spec.n_mlp):
        x = equal_linear(x, sd[f'style.{i + 1}.weight'], sd[f'style.{i + 1}.bias'], lr_mul=spec.lr_mlp,
                         activation=True)
    return x


def modulated_conv2d(sd, prefix, x, style, demodulate=True, upsample=False):
    """`ModulatedConv2d.forward`, model.py:237-278 (per-sample weights + grouped conv)."""
    weight_p = sd[f'{prefix}.weight']
    _, cout, cin, k, _ = weight_p.shape
    batch, in_channel, height, width = x.shape
    scale = 1 / math.sqrt(cin * k ** 2)
    style = equal_linear(style, sd[f'{prefix}.modulation.weight'], sd[f'{prefix}.modulation.bias'])
    style = style.view(batch, 1, in_channel, 1, 1)
    weight = scale * weight_p * style
    if demodulate:
        demod = torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8)
        weight = weight * demod.view(batch, cout, 1, 1, 1)
    weight = weight.view(batch * cout, in_channel, k, k)
    if upsample:
        x = x.view(1, batch * in_channel, height, width)
        weight = weight.view(batch, cout, in_channel, k, k)
        weight = weight.transpose(1, 2).reshape(batch * in_channel, cout, k, k)
        out = F.conv_transpose2d(x, weight, padding=0, stride=2, groups=batch)
        _, _, height, width = out.shape
        out = out.view(batch, cout, height, width)
        # Blur(pad=(pad0, pad1)) with factor 2, model.py:201-207: p = (4-2)-(3-1) = 0 -> pad (1, 1)
        kl = sd[f'{prefix}.blur.kernel']
        p = (kl.shape[0] - 2) - (k - 1)
        out = upfirdn2d(out, kl, pad=((p + 1) // 2 + 2 - 1, p // 2 + 1))
    else:
        x = x.view(1, batch * in_channel, height, width)
        out = F.conv2d(x, weight, padding=k // 2, groups=batch)
        _, _, height, width = out.shape
        out = out.view(batch, cout, height, width)
    return out


def styled_conv(sd, prefix, x, style, noise, upsample=False):
    """`StyledConv.forward`, model.py:336-342 (+ `NoiseInjection.forward` :287-292)."""
    out = modulated_conv2d(sd, f'{prefix}.conv', x, style, demodulate=True, upsample=upsample)
    if noise is None:
        b, _, h, w = out.shape
        noise = out.new_empty(b, 1, h, w).normal_()
    out = out + sd[f'{prefix}.noise.weight'] * noise
    return fused_leaky_relu(out, sd[f'{prefix}.activate.bias'])


def to_rgb(sd, prefix, x, style, skip=None):
    """`ToRGB.forward`, model.py:355-364 (+ `Upsample` :34-52: pad (2, 1), kernel*4)."""
    out = modulated_conv2d(sd, f'{prefix}.conv', x, style, demodulate=False)
    out = out + sd[f'{prefix}.bias']
    if skip is not None:
        kl = sd[f'{prefix}.upsample.kernel']
        p = kl.shape[0] - 2
        skip = upfirdn2d(skip, kl, up=2, down=1, pad=((p + 1) // 2 + 2 - 1, p // 2))
        out = out + skip
    return out


def make_noise(spec: GeneratorSpec) -> List[torch.Tensor]:
    """`Generator.make_noise`, model.py:457-466 (one 4x4, then two per resolution, shared over batch)."""
    noises = [torch.randn(1, 1, 4, 4)]
    for i in range(3, spec.log_size + 1):
        for _ in range(2):
            noises.append(torch.randn(1, 1, 2 ** i, 2 ** i))
    return noises


def mean_latent(sd, spec: GeneratorSpec, n_latent: int):
    """`Generator.mean_latent`, model.py:468-474."""
    latent_in = torch.randn(n_latent, spec.style_dim)
    return style_mlp(sd, spec, latent_in).mean(0, keepdim=True)


@torch.no_grad()
def generator_forward(sd, spec: GeneratorSpec, styles, return_latents=False, inject_index=None, truncation=1,
                      truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True,
                      return_intermediate_activations=False):
    """`Generator.forward`, model.py:479-561."""
    if not input_is_latent:
        styles = [style_mlp(sd, spec, s) for s in styles]
    if noise is None:
        if randomize_noise:
            noise = [None] * spec.num_layers
        else:
            noise = [sd[f'noises.noise_{i}'] for i in range(spec.num_layers)]
    if truncation < 1:
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
    if len(styles) < 2:
        inject_index = spec.n_latent
        if styles[0].ndim < 3:
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
        else:
            latent = styles[0]
    else:
        if inject_index is None:
            import random
            inject_index = random.randint(1, spec.n_latent - 1)
        latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
        latent2 = styles[1].unsqueeze(1).repeat(1, spec.n_latent - inject_index, 1)
        latent = torch.cat([latent, latent2], 1)

    acts = {} if return_intermediate_activations else None
    out = sd['input.input'].repeat(latent.shape[0], 1, 1, 1)
    if acts is not None:
        acts[0] = out.clone()
    out = styled_conv(sd, 'conv1', out, latent[:, 0], noise[0])
    if acts is not None:
        acts[1] = out.clone()
    skip = to_rgb(sd, 'to_rgb1', out, latent[:, 1])
    i = 1
    for j in range(spec.log_size - 2):
        out = styled_conv(sd, f'convs.{2 * j}', out, latent[:, i], noise[1 + 2 * j], upsample=True)
        if acts is not None:
            acts[i + 1] = out.clone()
        out = styled_conv(sd, f'convs.{2 * j + 1}', out, latent[:, i + 1], noise[2 + 2 * j])
        if acts is not None:
            acts[i + 2] = out.clone()
        skip = to_rgb(sd, f'to_rgbs.{j}', out, latent[:, i + 2], skip)
        i += 2
    image = skip
    if return_latents:
        return image, latent
    if return_intermediate_activations:
        return image, acts
    return image, None


def conv_flops_per_image(spec: GeneratorSpec) -> float:
    """Algorithmic generator FLOPs per image: 2*MACs with plain conv H^2*9*Cin*Cout,
    up-conv H_in^2*9*Cin*Cout, ToRGB H^2*Cin*3."""
    c4 = spec.channels[4]
    macs = 16 * 9 * c4 * c4 + 16 * c4 * 3
    cin = c4
    for i in range(3, spec.log_size + 1):
        cout = spec.channels[2 ** i]
        h_in, h = 2 ** (i - 1), 2 ** i
        macs += h_in * h_in * 9 * cin * cout + h * h * 9 * cout * cout + h * h * cout * 3
        cin = cout
    return 2.0 * macs
