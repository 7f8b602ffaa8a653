"""CPU oracle: SWAGAN generator forward with activation capture.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch-CPU fp32 restatement of
`scf/networks/swagan/model.py:14-286` on top of the StyleGAN2 oracle's layers (the reference's SWAGAN imports
ModulatedConv2d / StyledConv / ConstantInput / Upsample from the StyleGAN2 model, swagan/model.py:12): Haar wavelet
taps (:14-25), HaarTransform (:28-47), InverseHaarTransform (:49-68), the 12-channel ToRGB with its wavelet-domain skip
(:71-98) and Generator.forward (:211-286; the conv trunk stops at size / 2, the final image is the inverse transform of
the last skip).  State-dict keys are the reference's.

Pinned by `tests/golden/make_golden_swagan.py` against the reference's own `swagan/model.py` run in the build container
(bit-for-bit on CPU).
"""
import math
from collections import OrderedDict
from typing import List, Optional

import torch

from . import stylegan2_oracle as so


def get_haar_wavelet():
    """swagan/model.py:14-25: the four 2x2 Haar taps (ll, lh, hl, hh)."""
    l = 1 / (2 ** 0.5) * torch.ones(1, 2)
    h = 1 / (2 ** 0.5) * torch.ones(1, 2)
    h[0, 0] = -1 * h[0, 0]
    return l.T * l, h.T * l, l.T * h, h.T * h


def haar_transform(x, taps):
    """HaarTransform.forward, :39-45: four decimating 2x2 filters, concatenated along the channels."""
    return torch.cat([so.upfirdn2d(x, k, down=2) for k in taps], 1)


def inverse_haar_transform(x, taps):
    """InverseHaarTransform.forward, :61-68 (buffers ll, -lh, -hl, hh; up 2, pad (1, 0))."""
    ll, lh, hl, hh = x.chunk(4, 1)
    k_ll, k_lh, k_hl, k_hh = taps
    return (so.upfirdn2d(ll, k_ll, up=2, pad=(1, 0)) + so.upfirdn2d(lh, k_lh, up=2, pad=(1, 0))
            + so.upfirdn2d(hl, k_hl, up=2, pad=(1, 0)) + so.upfirdn2d(hh, k_hh, up=2, pad=(1, 0)))


class SwaganSpec:
    """Shape bookkeeping of swagan Generator.__init__, :101-176."""

    def __init__(self, size: int, style_dim: int, n_mlp: int, channel_multiplier: int = 2, blur_kernel=(1, 3, 3, 1), lr_mlp: float = 0.01):
        self.size, self.style_dim, self.n_mlp = size, style_dim, n_mlp
        self.channel_multiplier, self.blur_kernel, self.lr_mlp = channel_multiplier, list(blur_kernel), lr_mlp
        self.channels = so.get_channels(channel_multiplier)
        self.log_size = int(math.log(size, 2)) - 1
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.n_latent = self.log_size * 2 - 2


def init_state_dict(spec: SwaganSpec, seed: Optional[int] = None) -> 'OrderedDict[str, torch.Tensor]':
    """Random init in the reference's constructor order (:101-176): style MLP, constant input, conv1, to_rgb1 (12 output
    channels; Haar buffers only on the up-sampling ToRGBs), noise buffers, then (up conv, conv, to_rgb) per resolution,
    and the generator's own `iwt` buffers."""
    if seed is not None:
        torch.manual_seed(seed)
    sd = OrderedDict()
    sdim = spec.style_dim
    for i in range(spec.n_mlp):
        sd[f'style.{i + 1}.weight'] = torch.randn(sdim, sdim).div_(spec.lr_mlp)
        sd[f'style.{i + 1}.bias'] = torch.zeros(sdim)
    c4 = spec.channels[4]
    sd['input.input'] = torch.randn(1, c4, 4, 4)
    ll, lh, hl, hh = get_haar_wavelet()

    def modconv(prefix, cin, cout, k, with_blur):
        sd[f'{prefix}.weight'] = torch.randn(1, cout, cin, k, k)
        if with_blur:
            sd[f'{prefix}.blur.kernel'] = so.make_kernel(spec.blur_kernel) * 4
        sd[f'{prefix}.modulation.weight'] = torch.randn(cin, sdim)
        sd[f'{prefix}.modulation.bias'] = torch.ones(cin)

    def styled(prefix, cin, cout, up):
        modconv(f'{prefix}.conv', cin, cout, 3, up)
        sd[f'{prefix}.noise.weight'] = torch.zeros(1)
        sd[f'{prefix}.activate.bias'] = torch.zeros(cout)

    def torgb(prefix, cin, up):
        if up:
            sd[f'{prefix}.iwt.ll'], sd[f'{prefix}.iwt.lh'], sd[f'{prefix}.iwt.hl'], sd[f'{prefix}.iwt.hh'] = ll, -lh, -hl, hh
            sd[f'{prefix}.upsample.kernel'] = so.make_kernel(spec.blur_kernel) * 4
            sd[f'{prefix}.dwt.ll'], sd[f'{prefix}.dwt.lh'], sd[f'{prefix}.dwt.hl'], sd[f'{prefix}.dwt.hh'] = ll, lh, hl, hh
        modconv(f'{prefix}.conv', cin, 12, 1, False)
        sd[f'{prefix}.bias'] = torch.zeros(1, 12, 1, 1)

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
    sd['iwt.ll'], sd['iwt.lh'], sd['iwt.hl'], sd['iwt.hh'] = ll, -lh, -hl, hh
    return sd


def make_noise(spec: SwaganSpec) -> List[torch.Tensor]:
    """:178-187."""
    noises = [torch.randn(1, 1, 4, 4)]
    for i in range(3, spec.log_size + 1):
        for _ in range(2):
            noises.append(torch.randn(1, 1, 2 ** i, 2 ** i))
    return noises


def to_rgb(sd, prefix, x, style, skip=None):
    """swagan ToRGB.forward, :84-98: 1x1 modulated conv to 12 wavelet channels + bias; the skip goes image domain ->
    Upsample -> wavelet domain."""
    out = so.modulated_conv2d(sd, f'{prefix}.conv', x, style, demodulate=False)
    out = out + sd[f'{prefix}.bias']
    if skip is not None:
        skip = inverse_haar_transform(skip, [sd[f'{prefix}.iwt.{n}'] for n in ('ll', 'lh', 'hl', 'hh')])
        kl = sd[f'{prefix}.upsample.kernel']
        p = kl.shape[0] - 2
        skip = so.upfirdn2d(skip, kl, up=2, down=1, pad=((p + 1) // 2 + 2 - 1, p // 2))
        skip = haar_transform(skip, [sd[f'{prefix}.dwt.{n}'] for n in ('ll', 'lh', 'hl', 'hh')])
        out = out + skip
    return out


@torch.no_grad()
def generator_forward(sd, spec: SwaganSpec, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                      input_is_latent=False, noise=None, randomize_noise=True, return_intermediate_activations=False):
    """swagan Generator.forward, :200-286."""
    if not input_is_latent:
        styles = [so.style_mlp(sd, spec, s) for s in styles]
    if noise is None:
        noise = [None] * spec.num_layers if randomize_noise else [sd[f'noises.noise_{i}'] for i in range(spec.num_layers)]
    if truncation < 1:
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
    if len(styles) < 2:
        inject_index = spec.n_latent
        latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].ndim < 3 else styles[0]
    else:
        if inject_index is None:
            import random
            inject_index = random.randint(1, spec.n_latent - 1)
        latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                            styles[1].unsqueeze(1).repeat(1, spec.n_latent - inject_index, 1)], 1)
    acts = {} if return_intermediate_activations else None
    out = sd['input.input'].repeat(latent.shape[0], 1, 1, 1)
    if acts is not None:
        acts[0] = out.clone()
    out = so.styled_conv(sd, 'conv1', out, latent[:, 0], noise[0])
    if acts is not None:
        acts[1] = out.clone()
    skip = to_rgb(sd, 'to_rgb1', out, latent[:, 1])
    i = 1
    for j in range(spec.log_size - 2):
        out = so.styled_conv(sd, f'convs.{2 * j}', out, latent[:, i], noise[1 + 2 * j], upsample=True)
        if acts is not None:
            acts[i + 1] = out.clone()
        out = so.styled_conv(sd, f'convs.{2 * j + 1}', out, latent[:, i + 1], noise[2 + 2 * j])
        if acts is not None:
            acts[i + 2] = out.clone()
        skip = to_rgb(sd, f'to_rgbs.{j}', out, latent[:, i + 2], skip)
        i += 2
    image = inverse_haar_transform(skip, [sd[f'iwt.{n}'] for n in ('ll', 'lh', 'hl', 'hh')])
    if return_latents:
        return image, latent
    if return_intermediate_activations:
        return image, acts
    return image, None
