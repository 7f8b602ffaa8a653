"""GPU parity of the module-level forwards (the reference's sub-modules used outside Generator.forward) against the
oracle: EqualLinear, PixelNorm, ModulatedConv2d (plain / upsample / no demod), StyledConv, NoiseInjection, ToRGB,
Blur / Upsample buffers."""
import pytest
import torch

from oracle import stylegan2_oracle as so
from synthesis_in_style_b200 import model as M

pytestmark = pytest.mark.gpu


def sd_of(module, prefix):
    return {f'{prefix}.{k}': v.detach().cpu() for k, v in module.state_dict().items()}


def test_equal_linear_and_pixel_norm(cuda_device):
    torch.manual_seed(0)
    for act, lr in ((None, 1.0), ('fused_lrelu', 0.01)):
        lin = M.EqualLinear(96, 40, lr_mul=lr, activation=act, bias_init=0.3).to(cuda_device)
        x = torch.randn(7, 96)
        want = so.equal_linear(x, lin.weight.detach().cpu(), lin.bias.detach().cpu(), lr_mul=lr, activation=bool(act))
        got = lin(x.to(cuda_device)).cpu()
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    z = torch.randn(5, 512)
    torch.testing.assert_close(M.PixelNorm()(z.to(cuda_device)).cpu(), so.pixel_norm(z), rtol=1e-6, atol=1e-6)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        M.PixelNorm()(z)


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-5), ('bf16x3', 5e-4)])
@pytest.mark.parametrize('cfg', [(64, 128, 16, False, True), (128, 64, 8, True, True), (64, 64, 32, False, False), (32, 32, 16, True, True),
                                 # Cout = 128 with H a multiple of 32: the transposed-product kernels (plain: 32 x 8 halo tiles;
                                 # up: phase interiors + border strips), incl. a non-power-of-two size and demod off
                                 (128, 128, 32, False, True), (64, 128, 96, False, True), (96, 128, 64, False, False),
                                 (128, 128, 32, True, True), (64, 128, 64, True, True)])
def test_modulated_and_styled_conv(cuda_device, cfg, precision, tol):
    cin, cout, res, up, demod = cfg
    torch.manual_seed(cin + cout + res)
    conv = M.StyledConv(cin, cout, 3, 64, upsample=up, demodulate=demod).to(cuda_device)
    conv.conv.precision = precision
    with torch.no_grad():
        conv.noise.weight.fill_(0.37)
        conv.activate.bias.normal_(0, 0.1)
    sd = sd_of(conv, 'L')
    x, style = torch.randn(3, cin, res, res), torch.randn(3, 64)
    r = 2 * res if up else res
    noise = torch.randn(1, 1, r, r)
    want_mod = so.modulated_conv2d(sd, 'L.conv', x, style, demodulate=demod, upsample=up)
    want_sty = so.styled_conv(sd, 'L', x, style, noise, upsample=up) if demod else None
    with torch.no_grad():
        got_mod = conv.conv(x.to(cuda_device), style.to(cuda_device)).cpu()
        got_sty = conv(x.to(cuda_device), style.to(cuda_device), noise=noise.to(cuda_device)).cpu()
    scale = max(1.0, float(want_mod.abs().max()))
    assert float((got_mod - want_mod).abs().max()) <= tol * scale
    if want_sty is not None:
        assert float((got_sty - want_sty).abs().max()) <= tol * scale
    # per-sample noise and the random default
    with torch.no_grad():
        per = torch.randn(3, 1, r, r)
        got_ps = conv(x.to(cuda_device), style.to(cuda_device), noise=per.to(cuda_device)).cpu()
        assert conv(x.to(cuda_device), style.to(cuda_device)).shape == got_ps.shape
    if demod:
        want_ps = so.styled_conv(sd, 'L', x, style, per, upsample=up)
        assert float((got_ps - want_ps).abs().max()) <= tol * scale


def test_to_rgb_noise_injection_and_fir_modules(cuda_device):
    torch.manual_seed(3)
    rgb = M.ToRGB(128, 64).to(cuda_device)
    with torch.no_grad():
        rgb.bias.normal_(0, 0.1)
    sd = sd_of(rgb, 'R')
    x, style, skip = torch.randn(2, 128, 32, 32), torch.randn(2, 64), torch.randn(2, 3, 16, 16)
    want = so.to_rgb(sd, 'R', x, style, skip)
    want_noskip = so.to_rgb(sd, 'R', x, style, None)
    with torch.no_grad():
        got = rgb(x.to(cuda_device), style.to(cuda_device), skip.to(cuda_device)).cpu()
        got_noskip = rgb(x.to(cuda_device), style.to(cuda_device)).cpu()
    torch.testing.assert_close(got, want, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(got_noskip, want_noskip, rtol=1e-4, atol=2e-4)
    # Upsample / Blur buffers apply the reference's upfirdn2d configuration
    up = rgb.upsample(skip.to(cuda_device)).cpu()
    torch.testing.assert_close(up, so.upfirdn2d(skip, sd['R.upsample.kernel'], up=2, down=1, pad=(2, 1)), rtol=1e-5, atol=1e-5)
    conv = M.ModulatedConv2d(8, 8, 3, 16, upsample=True).to(cuda_device)
    t = torch.randn(1, 8, 17, 17)
    torch.testing.assert_close(conv.blur(t.to(cuda_device)).cpu(), so.upfirdn2d(t, conv.blur.kernel.cpu(), pad=(1, 1)), rtol=1e-5, atol=1e-5)
    # NoiseInjection and ConstantInput
    ni = M.NoiseInjection().to(cuda_device)
    with torch.no_grad():
        ni.weight.fill_(-0.6)
    img, nz = torch.randn(2, 5, 8, 8), torch.randn(1, 1, 8, 8)
    assert torch.equal(ni(img.to(cuda_device), nz.to(cuda_device)).cpu(), img + torch.tensor(-0.6) * nz)
    ci = M.ConstantInput(16).to(cuda_device)
    assert ci(torch.zeros(3, 4, device=cuda_device)).shape == (3, 16, 4, 4)
