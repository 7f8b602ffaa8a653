"""SWAGAN generator: CPU tests of the oracle against the golden vectors of the reference's own
`networks/swagan/model.py` and of the product's module tree / init order; GPU parity of the composed generator."""
import os

import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as so
from oracle import swagan_oracle as sw

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def gold():
    with np.load(os.path.join(HERE, 'golden', 'golden_swagan_v1.npz')) as z:
        return {k: z[k] for k in z.files}


def setup(gold):
    size, style_dim, n_mlp, batch = (int(v) for v in gold['cfg'])
    spec = sw.SwaganSpec(size, style_dim, n_mlp)
    sd = so.perturb_zero_params(sw.init_state_dict(spec, seed=0), seed=1234)
    z, z2, ml = torch.from_numpy(gold['z']), torch.from_numpy(gold['z2']), torch.from_numpy(gold['mean_latent'])
    noise = [torch.from_numpy(gold[f'noise/{i}']) for i in range(spec.num_layers)]
    cases = {'plain': dict(styles=[z]), 'truncmix': dict(styles=[z, z2], inject_index=2, truncation=0.7, truncation_latent=ml)}
    return spec, sd, noise, cases


def test_oracle_reproduces_reference_golden(gold):
    spec, sd, noise, cases = setup(gold)
    assert (spec.log_size, spec.num_layers, spec.n_latent) == (4, 5, 6)          # the trunk stops at size / 2
    for name, kw in cases.items():
        img, acts = sw.generator_forward(sd, spec, kw['styles'], noise=noise, return_intermediate_activations=True,
                                         **{k: v for k, v in kw.items() if k != 'styles'})
        assert torch.equal(img, torch.from_numpy(gold[f'{name}/image'])), name
        for k, a in acts.items():
            assert tuple(a.shape) == tuple(gold[f'{name}/act/{k}/shape'])
            np.testing.assert_array_equal(a.reshape(-1).numpy()[gold[f'{name}/act/{k}/idx']], gold[f'{name}/act/{k}/val'])
    x = torch.from_numpy(gold['haar/x'])
    taps = sw.get_haar_wavelet()
    w = sw.haar_transform(x, taps)
    assert torch.equal(w, torch.from_numpy(gold['haar/dwt']))
    assert torch.equal(sw.inverse_haar_transform(w, [taps[0], -taps[1], -taps[2], taps[3]]), torch.from_numpy(gold['haar/iwt']))
    torch.testing.assert_close(torch.from_numpy(gold['haar/iwt']), x, rtol=0, atol=1e-6)      # perfect reconstruction


def test_module_tree_and_init_order_match_the_reference(gold):
    from synthesis_in_style_b200 import swagan
    size, style_dim, n_mlp, _ = (int(v) for v in gold['cfg'])
    torch.manual_seed(0)
    g = swagan.Generator(size, style_dim, n_mlp)
    ref = sw.init_state_dict(sw.SwaganSpec(size, style_dim, n_mlp), seed=0)
    sd = g.state_dict()
    assert set(sd) == set(ref)
    for k in ref:
        assert torch.equal(sd[k], ref[k]), k        # same draw order: torch.manual_seed(s) gives the reference's weights
    assert (g.log_size, g.num_layers, g.n_latent) == (4, 5, 6)
    with torch.no_grad(), pytest.raises(RuntimeError, match='CUDA tensor'):
        g([torch.randn(1, style_dim)])


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_generator_matches_reference_golden_on_gpu(cuda_device, gold, precision):
    from synthesis_in_style_b200 import swagan
    spec, sd, noise, cases = setup(gold)
    g = swagan.Generator(spec.size, spec.style_dim, spec.n_mlp, precision=precision)
    g.load_state_dict(sd)
    g = g.to(cuda_device).eval()
    tol = 2e-4 if precision == 'fp32' else 2e-3
    for name, kw in cases.items():
        args = {k: (v.to(cuda_device) if isinstance(v, torch.Tensor) else v) for k, v in kw.items() if k != 'styles'}
        with torch.no_grad():
            img, acts = g([s.to(cuda_device) for s in kw['styles']], noise=[n.to(cuda_device) for n in noise],
                          return_intermediate_activations=True, **args)
        want = torch.from_numpy(gold[f'{name}/image'])
        assert img.shape == want.shape == (2, 3, spec.size, spec.size)
        assert float((img.cpu() - want).abs().max()) <= tol * max(1.0, float(want.abs().max())), name
        assert sorted(acts) == list(range(spec.n_latent))
        for k, a in acts.items():
            got = a.reshape(-1).cpu().numpy()[gold[f'{name}/act/{k}/idx']]
            ref = gold[f'{name}/act/{k}/val']
            assert np.abs(got - ref).max() <= tol * max(1.0, float(np.abs(ref).max())), (name, k)
    # the wavelet modules alone: bit-exact against the reference's CPU result is not required (FMA order), 1e-6 is
    x = torch.from_numpy(gold['haar/x']).to(cuda_device)
    dwt, iwt = swagan.HaarTransform(3).to(cuda_device), swagan.InverseHaarTransform(3).to(cuda_device)
    torch.testing.assert_close(dwt(x).cpu(), torch.from_numpy(gold['haar/dwt']), rtol=0, atol=1e-6)
    torch.testing.assert_close(iwt(dwt(x)).cpu(), torch.from_numpy(gold['haar/iwt']), rtol=0, atol=1e-6)
