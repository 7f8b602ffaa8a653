"""GPU parity: Generator.forward through the C-ABI against the golden vectors (reference outputs) and the oracle.

Tolerances (north star): image max abs error <= 2e-2 and PSNR >= 40 dB (peak 2.0) against the fp32 reference;
the fp32 CUDA-core path and the bf16x3 tensor-core path are both held to a much tighter 2e-3 here."""
import math

import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as so
from synthesis_in_style_b200.model import Generator

pytestmark = pytest.mark.gpu

IMG_ATOL = 2e-2
TIGHT = {'fp32': 2e-4, 'bf16x3': 2e-3}


def psnr(a, b, peak=2.0):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else 10 * math.log10(peak * peak / mse)


def build(golden, tag, device, precision):
    size, sdim, n_mlp, cm, batch = [int(v) for v in golden[f'{tag}/config']]
    spec = so.GeneratorSpec(size, sdim, n_mlp, cm)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    torch.manual_seed(0)
    g = Generator(size, sdim, n_mlp, channel_multiplier=cm, precision=precision)
    g.load_state_dict(sd)
    return spec, sd, g.to(device).eval(), batch


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('tag', ['g32', 'g16trunc', 'g16mix'])
def test_generator_matches_reference_golden(golden, cuda_device, tag, precision):
    spec, sd, g, batch = build(golden, tag, cuda_device, precision)
    torch.manual_seed(1)
    z = torch.randn(batch, spec.style_dim)
    noise = [n.to(cuda_device) for n in so.make_noise(spec)]
    styles = [z.to(cuda_device)]
    kwargs = dict(noise=noise, return_intermediate_activations=True)
    trunc = float(golden[f'{tag}/truncation'][0])
    if trunc < 1:
        kwargs.update(truncation=trunc, truncation_latent=torch.from_numpy(golden[f'{tag}/mean_latent']).to(cuda_device))
    if f'{tag}/z_b' in golden:
        styles.append(torch.from_numpy(golden[f'{tag}/z_b']).to(cuda_device))
        kwargs.update(inject_index=3)
    with torch.no_grad():
        img, acts = g(styles, **kwargs)
        img2, lat = g(styles, **{**kwargs, 'return_intermediate_activations': False, 'return_latents': True,
                                'noise': None, 'randomize_noise': False})
    want = torch.from_numpy(golden[f'{tag}/image'])
    err = float((img.cpu() - want).abs().max())
    assert err <= IMG_ATOL and psnr(img.cpu().clamp(-1, 1), want.clamp(-1, 1)) >= 40.0
    assert err <= TIGHT[precision] * float(want.abs().max()), err
    assert sorted(acts) == list(range(spec.n_latent))
    for k, a in acts.items():
        assert list(a.shape) == [int(v) for v in golden[f'{tag}/act{k}/shape']]
        idx = torch.from_numpy(golden[f'{tag}/act{k}/idx'])
        got = a.reshape(-1)[idx.to(cuda_device)].cpu().numpy()
        val = golden[f'{tag}/act{k}/val']
        assert np.abs(got - val).max() <= TIGHT[precision] * max(1.0, np.abs(val).max()), (k, np.abs(got - val).max())
        np.testing.assert_allclose(float(a.double().abs().sum()), golden[f'{tag}/act{k}/sum_abs'][0], rtol=1e-4)
    np.testing.assert_allclose(lat.cpu().numpy(), golden[f'{tag}/latent'], rtol=0, atol=2e-5)
    assert float((img2.cpu() - torch.from_numpy(golden[f'{tag}/image_buffer_noise'])).abs().max()) <= TIGHT[precision] * float(want.abs().max())


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_generator_256_against_oracle(cuda_device, precision):
    """BASELINE config (256^2, 512-d styles, 8-layer MLP) on 2 samples against the CPU oracle."""
    spec = so.GeneratorSpec(256, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    g = Generator(256, 512, 8, precision=precision)
    g.load_state_dict(sd)
    g = g.to(cuda_device).eval()
    torch.manual_seed(1)
    z = torch.randn(2, 512)
    noise = so.make_noise(spec)
    want_img, want_acts = so.generator_forward(sd, spec, [z], noise=noise, return_intermediate_activations=True)
    with torch.no_grad():
        img, acts = g([z.to(cuda_device)], noise=[n.to(cuda_device) for n in noise], return_intermediate_activations=True)
    err = float((img.cpu() - want_img).abs().max())
    assert err <= IMG_ATOL, err
    assert psnr(img.cpu().clamp(-1, 1), want_img.clamp(-1, 1)) >= 40.0
    for k in want_acts:
        e = float((acts[k].cpu() - want_acts[k]).abs().max())
        assert e <= TIGHT[precision] * max(1.0, float(want_acts[k].abs().max())), (k, e)
    # capture_layers restricts what is materialised but not what is computed
    with torch.no_grad():
        img3, acts3 = g([z.to(cuda_device)], noise=[n.to(cuda_device) for n in noise], return_intermediate_activations=True,
                        capture_layers=[0, 8, 13])
    assert sorted(acts3) == [0, 8, 13] and torch.equal(img3, img) and torch.equal(acts3[13], acts[13])


def test_generator_api_surface(cuda_device):
    g = Generator(16, 64, 2).to(cuda_device).eval()
    z = torch.randn(3, 64, device=cuda_device)
    with torch.no_grad():
        w = g.get_latent(z)
        assert w.shape == (3, 64)
        torch.manual_seed(5)
        ml = g.mean_latent(128)
        assert ml.shape == (1, 64)
        # W+ input is used as-is (model.py:515-519)
        img_a, lat = g([z], return_latents=True, randomize_noise=False)
        img_b, _ = g([lat], input_is_latent=True, randomize_noise=False)
        assert torch.equal(img_a, img_b)
        img_c, _ = g([w], input_is_latent=True, randomize_noise=False)
        assert torch.equal(img_a, img_c)
        # randomize_noise=True draws per-sample noise in layer order from the device generator
        torch.manual_seed(9)
        img_r1, _ = g([z])
        torch.manual_seed(9)
        per_sample = [torch.empty(3, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2), device=cuda_device).normal_()
                      for l in range(g.num_layers)]
        img_r2, _ = g([z], noise=per_sample)
        assert torch.equal(img_r1, img_r2)
        out, none = g([z], randomize_noise=False)
        assert none is None and out.shape == (3, 3, 16, 16)
        empty, _ = g([z[:0]], randomize_noise=False)
        assert empty.shape == (0, 3, 16, 16)
    with pytest.raises(RuntimeError, match='inference-only'):
        g([z])
    # in-place weight updates are picked up (plan is re-prepared)
    with torch.no_grad():
        before, _ = g([z], randomize_noise=False)
        g.to_rgbs[1].bias.add_(0.5)
        after, _ = g([z], randomize_noise=False)
    torch.testing.assert_close(after, before + 0.5, rtol=0, atol=1e-5)


def _labels_agree(acts, want_acts, layers, k=4, seed=5):
    """north-star label criterion on the given layers with seeded unit-norm centroids."""
    from oracle import labelling_oracle as lo
    from synthesis_in_style_b200 import labelling
    g = torch.Generator().manual_seed(seed)
    total = agree = 0
    for layer in layers:
        c = want_acts[layer].shape[1]
        cent = torch.nn.functional.normalize(torch.randn(k, c, generator=g), dim=1)
        want, margin = lo.predict_with_margin(want_acts[layer], cent)
        got = labelling.FactorCatalog(k, cent).predict(acts[layer]).cpu()
        safe = margin > 1e-3
        assert torch.equal(got[safe], want[safe]), (layer, int((got[safe] != want[safe]).sum()))
        total += want.numel(); agree += int((got == want).sum())
    assert agree / total >= 0.999


def test_generator_512_config3(cuda_device):
    """BASELINE config 3 (512^2 config-f, multi-layer capture of the 64..512 px maps) at batch 2 against the oracle."""
    spec = so.GeneratorSpec(512, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    g = Generator(512, 512, 8)
    g.load_state_dict(sd)
    g = g.to(cuda_device).eval()
    torch.manual_seed(1)
    z = torch.randn(2, 512)
    noise = so.make_noise(spec)
    want_img, want_acts = so.generator_forward(sd, spec, [z], noise=noise, return_intermediate_activations=True)
    layers = list(range(8, 16))
    with torch.no_grad():
        img, acts = g([z.to(cuda_device)], noise=[n.to(cuda_device) for n in noise], return_intermediate_activations=True,
                      capture_layers=[0] + layers)
    assert sorted(acts) == [0] + layers
    assert float((img.cpu() - want_img).abs().max()) <= IMG_ATOL
    assert psnr(img.cpu().clamp(-1, 1), want_img.clamp(-1, 1)) >= 40.0
    for k in layers:
        e = float((acts[k].cpu() - want_acts[k]).abs().max())
        assert e <= TIGHT['bf16x3'] * max(1.0, float(want_acts[k].abs().max())), (k, e)
    _labels_agree(acts, want_acts, [8, 9, 14, 15])


def test_generator_1024_config4(cuda_device):
    """BASELINE config 4 (1024^2, truncation 0.7 with an injected mean latent, style mixing with an explicit
    inject_index) at batch 1 against the oracle."""
    spec = so.GeneratorSpec(1024, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    g = Generator(1024, 512, 8)
    g.load_state_dict(sd)
    g = g.to(cuda_device).eval()
    torch.manual_seed(7)
    ml = so.mean_latent(sd, spec, 256)
    torch.manual_seed(1)
    za, zb = torch.randn(1, 512), torch.randn(1, 512)
    noise = so.make_noise(spec)
    kw = dict(truncation=0.7, inject_index=6, return_intermediate_activations=True)
    want_img, want_acts = so.generator_forward(sd, spec, [za, zb], noise=noise, truncation_latent=ml, **kw)
    with torch.no_grad():
        torch.manual_seed(7)
        ml_gpu = g.mean_latent(256)       # same CPU->device semantics are NOT guaranteed (device RNG); compare loosely
        img, acts = g([za.to(cuda_device), zb.to(cuda_device)], noise=[n.to(cuda_device) for n in noise],
                      truncation_latent=ml.to(cuda_device), capture_layers=[0, 16, 17], **kw)
    assert ml_gpu.shape == (1, 512) and torch.isfinite(ml_gpu).all()
    assert float((img.cpu() - want_img).abs().max()) <= IMG_ATOL
    assert psnr(img.cpu().clamp(-1, 1), want_img.clamp(-1, 1)) >= 40.0
    for k in (16, 17):
        e = float((acts[k].cpu() - want_acts[k]).abs().max())
        assert e <= TIGHT['bf16x3'] * max(1.0, float(want_acts[k].abs().max())), (k, e)
    _labels_agree(acts, want_acts, [16, 17])
