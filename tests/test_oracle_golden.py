"""CPU: the oracle against the golden vectors produced by the reference itself (tests/golden/make_golden.py)."""
import json

import numpy as np
import pytest
import torch

from oracle import labelling_oracle as lo
from oracle import stylegan2_oracle as so

UPFIRDN_CASES = ['blur_mode1', 'up_mode3', 'down_mode5', 'asym_mode1', 'k3_mode2', 'haar_up_mode4', 'haar_down_mode6',
                 'negpad', 'minor3', 'up3_generic']


@pytest.mark.parametrize('name', UPFIRDN_CASES)
def test_upfirdn2d_oracle_matches_reference_native(golden, name):
    x = torch.from_numpy(golden[f'upfirdn2d/{name}/x'])
    k = torch.from_numpy(golden[f'upfirdn2d/{name}/k'])
    args = [int(v) for v in golden[f'upfirdn2d/{name}/args']]
    y = so.upfirdn2d_op(x, k, *args)
    assert torch.equal(y, torch.from_numpy(golden[f'upfirdn2d/{name}/y']))
    # and the restated CUDA-kernel index math agrees (tap flip, floor_div, phase)
    emu = so.upfirdn2d_index_emulation(x.numpy(), k.numpy(), *args)
    np.testing.assert_allclose(emu, golden[f'upfirdn2d/{name}/y'], atol=1e-5)


def test_fused_leaky_relu_oracle(golden):
    y = so.fused_leaky_relu(torch.from_numpy(golden['fused/x']), torch.from_numpy(golden['fused/b']))
    assert torch.equal(y, torch.from_numpy(golden['fused/y']))


def test_fused_bias_act_modes():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 4, generator=g)
    b = torch.randn(3, generator=g)
    ref = torch.randn(2, 3, 4, generator=g)
    e = x.new_empty(0)
    xb = x + b.view(1, 3, 1)
    assert torch.equal(so.fused_bias_act(x, b, e, 1, 0, 0.2, 2.0), xb * 2.0)
    assert torch.equal(so.fused_bias_act(x, b, e, 3, 2, 0.2, 2.0), torch.zeros_like(x))
    assert torch.equal(so.fused_bias_act(x, e, ref, 3, 1, 0.2, 1.5), torch.where(ref > 0, x, x * 0.2) * 1.5)


def _run_generator(golden, tag):
    size, sdim, n_mlp, cm, batch = [int(v) for v in golden[f'{tag}/config']]
    spec = so.GeneratorSpec(size, sdim, n_mlp, cm)
    sd = so.init_state_dict(spec, seed=0)
    for key in ('style.1.weight', 'conv1.conv.weight', f'convs.{2 * (spec.log_size - 3)}.conv.weight', 'noises.noise_2'):
        idx = golden[f'{tag}/w/{key}/idx']
        np.testing.assert_array_equal(sd[key].reshape(-1).numpy()[idx], golden[f'{tag}/w/{key}/val'])
    so.perturb_zero_params(sd, seed=1234)
    torch.manual_seed(1)
    z = torch.randn(batch, sdim)
    noise = so.make_noise(spec)
    np.testing.assert_array_equal(z.numpy(), golden[f'{tag}/z'])
    kwargs = dict(noise=noise, return_intermediate_activations=True)
    trunc = float(golden[f'{tag}/truncation'][0])
    styles = [z]
    if trunc < 1:
        kwargs.update(truncation=trunc, truncation_latent=torch.from_numpy(golden[f'{tag}/mean_latent']))
    if f'{tag}/z_b' in golden:
        styles = [z, torch.from_numpy(golden[f'{tag}/z_b'])]
        kwargs.update(inject_index=3)
    img, acts = so.generator_forward(sd, spec, styles, **kwargs)
    return spec, sd, img, acts


@pytest.mark.parametrize('tag', ['g32', 'g16trunc', 'g16mix'])
def test_generator_oracle_matches_reference(golden, tag):
    spec, sd, img, acts = _run_generator(golden, tag)
    # same ATen CPU ops in the same order as the reference: equal up to thread-count dependent summation order
    np.testing.assert_allclose(img.numpy(), golden[f'{tag}/image'], rtol=0, atol=2e-5)
    assert sorted(acts) == list(range(spec.n_latent))
    for k, a in acts.items():
        assert list(a.shape) == [int(v) for v in golden[f'{tag}/act{k}/shape']]
        idx = golden[f'{tag}/act{k}/idx']
        np.testing.assert_allclose(a.reshape(-1).numpy()[idx], golden[f'{tag}/act{k}/val'], rtol=0, atol=2e-5)
        np.testing.assert_allclose(np.abs(a.numpy().astype(np.float64)).sum(), golden[f'{tag}/act{k}/sum_abs'][0], rtol=1e-5)


def test_mean_latent_and_flops(golden):
    spec = so.GeneratorSpec(16, 64, 2, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    torch.manual_seed(7)
    ml = so.mean_latent(sd, spec, 64)
    np.testing.assert_allclose(ml.numpy(), golden['g16trunc/mean_latent'], atol=1e-6)
    # SURVEY.md §8d / BASELINE.md §2: 90.24 / 119.33 / 148.52 GFLOP per image
    for size, gf in ((256, 90.24), (512, 119.33), (1024, 148.52)):
        assert abs(so.conv_flops_per_image(so.GeneratorSpec(size, 512, 8, 2)) / 1e9 - gf) < 0.01


def test_labelling_oracle_matches_reference(golden):
    spec, sd, img, acts = _run_generator(golden, 'g32')
    raw_map = json.loads(bytes(golden['label/class_map_json']).decode())
    inv = lo.invert_class_label_map(raw_map)
    catalog = {layer: torch.from_numpy(golden[f'label/{layer}/centroids']) for layer in raw_map}
    total = agree = 0
    for layer, cent in catalog.items():
        ids, margin = lo.predict_with_margin(acts[int(layer)], cent)
        want = torch.from_numpy(golden[f'label/{layer}/ids']).long()
        gm = torch.from_numpy(golden[f'label/{layer}/margin'])
        safe = gm > 1e-3
        assert torch.equal(ids[safe], want[safe])
        total += ids.numel()
        agree += int((ids == want).sum())
    assert agree / total >= 0.999
    pc = lo.prepare_image_segmentation(acts, catalog, inv, 32)
    pc = lo.merge_sub_images(pc, {'merged': ['4', '6']}, ['background', 'printed_text', 'handwritten_text'])
    n = mism = 0
    for layer in pc:
        for cn, m in pc[layer].items():
            want = np.unpackbits(golden[f'label/mask/{layer}/{cn}'])[:m.numel()].reshape(m.shape).astype(bool)
            mism += int((m.numpy() != want).sum())
            n += m.numel()
    assert mism / n <= 1e-3
    ids_b = lo.bilinear_then_predict(acts[4], catalog['4'], 32)
    assert (ids_b.numpy() == golden['label/4/ids_bilinear']).mean() >= 0.999


def test_make_image_restatement():
    x = torch.tensor([[-2.0, -1.0, -0.5, 0.0], [0.25, 0.999, 1.0, 3.0]]).view(1, 1, 2, 4).repeat(1, 3, 1, 1)
    out = lo.make_image(x)
    assert out.shape == (1, 2, 4, 3) and out.dtype == torch.uint8
    assert out[0, :, :, 0].tolist() == [[0, 0, 63, 127], [159, 254, 255, 255]]
