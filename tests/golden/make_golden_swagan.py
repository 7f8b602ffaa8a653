"""Golden vectors of the SWAGAN generator, produced by running the REFERENCE's own `networks/swagan/model.py` on CPU.

Run in the build container only (`python tests/golden/make_golden_swagan.py`).  The reference module is imported in
isolation (synthetic package `refnet` with sub-packages `swagan` and `stylegan2`, the op packages' JIT build replaced by a
stub, `fused_leaky_relu` / `upfirdn2d` bound to the CPU forms as in make_golden.py).  The script ASSERTS that
`oracle/swagan_oracle.py` reproduces the reference bit-for-bit (weights from the same seed, image, every captured
activation, with and without truncation + style mixing) and then writes `golden_swagan_v1.npz`: inputs plus sampled
outputs (the state dict is re-derived from the seed by the tests).
"""
import importlib
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path('/root/reference/stylegan_code_finder')

from oracle import stylegan2_oracle as so  # noqa: E402
from oracle import swagan_oracle as sw  # noqa: E402


def import_reference_swagan():
    net = types.ModuleType('refnet')
    net.__path__ = [str(REF / 'networks')]
    sys.modules['refnet'] = net
    for sub in ('swagan', 'stylegan2'):
        pkg = types.ModuleType(f'refnet.{sub}')
        pkg.__path__ = [str(REF / 'networks' / sub)]
        sys.modules[f'refnet.{sub}'] = pkg
    import torch.utils.cpp_extension as cpp_ext
    real_load = cpp_ext.load
    cpp_ext.load = lambda *a, **k: types.SimpleNamespace()
    try:
        mods = {}
        for sub in ('swagan', 'stylegan2'):
            op_pkg = types.ModuleType(f'refnet.{sub}.op')
            op_pkg.__path__ = [str(REF / 'networks' / sub / 'op')]
            sys.modules[f'refnet.{sub}.op'] = op_pkg
            up = importlib.import_module(f'refnet.{sub}.op.upfirdn2d')
            fa = importlib.import_module(f'refnet.{sub}.op.fused_act')
            up.F = F
            mods[sub] = (op_pkg, up, fa)
    finally:
        cpp_ext.load = real_load

    def lrelu(x, b, negative_slope=0.2, scale=2 ** 0.5):
        return F.leaky_relu(x + b.view(1, b.shape[0], *[1] * (x.ndim - 2)), negative_slope) * scale

    def make_upfirdn(up):
        def upfirdn(x, k, up_=1, down=1, pad=(0, 0), **kw):
            up_f = kw.get('up', up_)
            b, c, h, w = x.shape
            o = up.upfirdn2d_native(x.reshape(-1, h, w, 1), k, up_f, up_f, down, down, pad[0], pad[1], pad[0], pad[1])
            return o.view(-1, c, o.shape[1], o.shape[2])
        return lambda x, k, up=1, down=1, pad=(0, 0): upfirdn(x, k, up, down, pad)

    for sub, (op_pkg, up, fa) in mods.items():
        class FusedLeakyReLU(torch.nn.Module):
            def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
                super().__init__()
                self.bias = torch.nn.Parameter(torch.zeros(channel))
                self.negative_slope, self.scale = negative_slope, scale

            def forward(self, x):
                return lrelu(x, self.bias, self.negative_slope, self.scale)
        op_pkg.FusedLeakyReLU = FusedLeakyReLU
        op_pkg.fused_leaky_relu = lrelu
        op_pkg.upfirdn2d = make_upfirdn(up)
        op_pkg.conv2d_gradfix = types.SimpleNamespace()
    sg2 = importlib.import_module('refnet.stylegan2.model')
    sg2.fused_leaky_relu = lrelu
    model = importlib.import_module('refnet.swagan.model')
    return model


def eq(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), f'{what}: oracle != reference, max|diff|={(a.float() - b.float()).abs().max().item()}'


def sample_idx(numel, n=512, seed=0):
    g = np.random.RandomState(seed)
    return np.sort(g.choice(numel, size=min(n, numel), replace=False)).astype(np.int64)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model = import_reference_swagan()
    out = {}
    size, style_dim, n_mlp, batch = 32, 64, 2, 2
    torch.manual_seed(0)
    ref_g = model.Generator(size, style_dim, n_mlp).eval()
    spec = sw.SwaganSpec(size, style_dim, n_mlp)
    sd = sw.init_state_dict(spec, seed=0)
    ref_sd = ref_g.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), sorted(set(ref_sd) ^ set(sd))
    for k in ref_sd:
        eq(sd[k], ref_sd[k], f'init {k}')
    sd = so.perturb_zero_params(sd, seed=1234)
    ref_g.load_state_dict(sd)
    assert (ref_g.log_size, ref_g.num_layers, ref_g.n_latent) == (spec.log_size, spec.num_layers, spec.n_latent)
    torch.manual_seed(1)
    z, z2 = torch.randn(batch, style_dim), torch.randn(batch, style_dim)
    noise = sw.make_noise(spec)
    mean_latent = so.style_mlp(sd, spec, torch.randn(64, style_dim)).mean(0, keepdim=True)
    cases = {'plain': dict(styles=[z]), 'truncmix': dict(styles=[z, z2], inject_index=2, truncation=0.7, truncation_latent=mean_latent)}
    out['cfg'] = np.array([size, style_dim, n_mlp, batch])
    out['z'], out['z2'], out['mean_latent'] = z.numpy(), z2.numpy(), mean_latent.numpy()
    for i, n in enumerate(noise):
        out[f'noise/{i}'] = n.numpy()
    for name, kw in cases.items():
        with torch.no_grad():
            ref_img, ref_acts = ref_g(kw['styles'], noise=noise, return_intermediate_activations=True,
                                      **{k: v for k, v in kw.items() if k != 'styles'})
        img, acts = sw.generator_forward(sd, spec, kw['styles'], noise=noise, return_intermediate_activations=True,
                                         **{k: v for k, v in kw.items() if k != 'styles'})
        eq(img, ref_img, f'{name} image')
        assert sorted(acts) == sorted(ref_acts)
        for k in acts:
            eq(acts[k], ref_acts[k], f'{name} act {k}')
        out[f'{name}/image'] = ref_img.numpy()
        for k, a in ref_acts.items():
            idx = sample_idx(a.numel(), seed=k)
            out[f'{name}/act/{k}/idx'], out[f'{name}/act/{k}/val'], out[f'{name}/act/{k}/shape'] = idx, a.reshape(-1).numpy()[idx], np.array(a.shape)
    # the wavelet transforms on their own (round trip is the identity up to rounding: pin the actual numbers)
    x = torch.randn(2, 3, 8, 8)
    taps = sw.get_haar_wavelet()
    ref_dwt, ref_iwt = model.HaarTransform(3), model.InverseHaarTransform(3)
    eq(sw.haar_transform(x, taps), ref_dwt(x), 'dwt')
    w = ref_dwt(x)
    eq(sw.inverse_haar_transform(w, [taps[0], -taps[1], -taps[2], taps[3]]), ref_iwt(w), 'iwt')
    out['haar/x'], out['haar/dwt'], out['haar/iwt'] = x.numpy(), ref_dwt(x).numpy(), ref_iwt(w).numpy()
    path = HERE / 'golden_swagan_v1.npz'
    np.savez_compressed(path, **out)
    print(f'wrote {path} ({path.stat().st_size / 1024:.0f} KiB); oracle == reference on {len(cases)} generator cases + Haar transforms')


if __name__ == '__main__':
    main()
