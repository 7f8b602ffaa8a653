"""Generate golden vectors by running the REFERENCE ITSELF (from /root/reference) on CPU.

Run in the build container only (`python tests/golden/make_golden.py`); /root/reference does not
exist on the GPU box, so tests read the committed `.npz` files, never the reference.

What is executed from the reference, unmodified:
  * `networks/stylegan2/model.py`  (Generator and all sub-modules), imported in isolation through a
    synthetic package so that `networks/__init__.py`'s heavy imports are bypassed;
  * `networks/stylegan2/op/upfirdn2d.py::upfirdn2d_native`  (with the `F` import it forgot);
  * `segmentation/gan_local_edit/factor_catalog.py::FactorCatalog.predict / pairwise_distance`
    (sklearn-private import stubbed; `.cuda()` made a no-op because this container has no GPU);
  * `segmentation/base_cluster_based_dataset_segmenter.py::predict_clusters`,
    `segmentation/base_dataset_segmenter.py::resize_to_image_size`,
    `segmentation/black_white_handwritten_printed_text_segmenter.py::merge_sub_images`.
The only restated piece is `fused_leaky_relu` (the reference has no CPU version of it at all; restated
from fused_bias_act_kernel.cu:26-47 as leaky_relu(x+b, 0.2)*sqrt(2)).

The script ASSERTS that the oracle (`oracle/`) reproduces every reference output bit-for-bit, then
writes the vectors.
"""
import importlib
import importlib.machinery
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
from oracle import labelling_oracle as lo  # noqa: E402


def import_reference_generator():
    """SURVEY.md Appendix B2/B3."""
    op_pkg = types.ModuleType('refsg2.op')
    op_pkg.__path__ = [str(REF / 'networks/stylegan2/op')]
    pkg = types.ModuleType('refsg2')
    pkg.__path__ = [str(REF / 'networks/stylegan2')]
    sys.modules['refsg2'] = pkg

    # The op package's __init__ JIT-builds CUDA extensions at import; replace torch's `load` by a stub so
    # the python files import (their CUDA entry points are never called on CPU).
    import torch.utils.cpp_extension as cpp_ext
    real_load = cpp_ext.load
    cpp_ext.load = lambda *a, **k: types.SimpleNamespace()
    try:
        upmod = importlib.import_module('refsg2.op.upfirdn2d')
        fa = importlib.import_module('refsg2.op.fused_act')
    finally:
        cpp_ext.load = real_load
    upmod.F = F  # the reference forgot this import (op/upfirdn2d.py:1-5)

    def lrelu(x, b, negative_slope=0.2, scale=2 ** 0.5):
        return F.leaky_relu(x + b.view(1, b.shape[0], *[1] * (x.ndim - 2)), negative_slope) * scale

    def upfirdn(x, k, up=1, down=1, pad=(0, 0)):
        b, c, h, w = x.shape
        o = upmod.upfirdn2d_native(x.reshape(-1, h, w, 1), k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        return o.view(-1, c, o.shape[1], o.shape[2])

    fa.fused_leaky_relu = lrelu
    opmod = importlib.import_module('refsg2.op')
    opmod.fused_leaky_relu = lrelu
    opmod.upfirdn2d = upfirdn
    model = importlib.import_module('refsg2.model')
    model.fused_leaky_relu = lrelu
    model.upfirdn2d = upfirdn
    return model, upmod


def import_reference_labelling():
    sys.path.insert(0, str(REF))
    stub = types.ModuleType('segmentation.gan_local_edit.spherical_kmeans')

    class MiniBatchSphericalKMeans:  # only the attribute used at inference
        def __init__(self, n_clusters=0, random_state=0, **kw):
            self.cluster_centers_ = None
    stub.MiniBatchSphericalKMeans = MiniBatchSphericalKMeans
    sys.modules['segmentation.gan_local_edit.spherical_kmeans'] = stub
    fc = importlib.import_module('segmentation.gan_local_edit.factor_catalog')
    base = importlib.import_module('segmentation.base_cluster_based_dataset_segmenter')
    bw = importlib.import_module('segmentation.black_white_handwritten_printed_text_segmenter')
    return fc, base, bw


def eq(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), f'{what}: oracle != reference, max|diff|={(a.float() - b.float()).abs().max().item()}'


def sample_idx(numel, n=512, seed=0):
    g = np.random.RandomState(seed)
    return np.sort(g.choice(numel, size=min(n, numel), replace=False)).astype(np.int64)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)
    model, upmod = import_reference_generator()
    out = {}

    # ------------------------------------------------------------------ ops
    g = torch.Generator().manual_seed(11)
    cases = []
    k4 = so.make_kernel([1, 3, 3, 1])
    kasym = torch.randn(4, 4, generator=g)
    k3 = torch.randn(3, 3, generator=g)
    k2 = torch.tensor([[0.5, -0.5], [0.25, 1.0]])
    k43 = torch.randn(4, 3, generator=g)
    # (name, shape[major,H,W,minor], kernel, up, down, pad(x0,x1,y0,y1))
    cases.append(('blur_mode1', (6, 17, 17, 1), k4 * 4, 1, 1, (1, 1, 1, 1)))
    cases.append(('up_mode3', (6, 8, 8, 1), k4 * 4, 2, 1, (2, 1, 2, 1)))
    cases.append(('down_mode5', (4, 16, 16, 1), k4, 1, 2, (1, 1, 1, 1)))
    cases.append(('asym_mode1', (3, 9, 13, 1), kasym, 1, 1, (2, 1, 0, 3)))
    cases.append(('k3_mode2', (3, 10, 7, 1), k3, 1, 1, (1, 1, 1, 1)))
    cases.append(('haar_up_mode4', (3, 6, 6, 1), k2, 2, 1, (1, 0, 1, 0)))
    cases.append(('haar_down_mode6', (3, 8, 8, 1), k2, 1, 2, (0, 0, 0, 0)))
    cases.append(('negpad', (2, 12, 12, 1), kasym, 1, 1, (-1, 2, 1, -2)))
    cases.append(('minor3', (2, 7, 9, 3), k43, 2, 1, (2, 1, 1, 1)))
    cases.append(('up3_generic', (2, 5, 6, 1), kasym, 3, 2, (2, 2, 3, 1)))
    for name, shape, k, up, down, pad in cases:
        x = torch.randn(*shape, generator=g)
        ref = upmod.upfirdn2d_native(x, k, up, up, down, down, *pad)
        mine = so.upfirdn2d_op(x, k, up, up, down, down, *pad)
        eq(mine, ref, f'upfirdn2d {name}')
        emu = so.upfirdn2d_index_emulation(x.numpy(), k.numpy(), up, up, down, down, *pad)
        assert np.abs(emu - ref.numpy()).max() < 1e-5, name
        out[f'upfirdn2d/{name}/x'] = x.numpy()
        out[f'upfirdn2d/{name}/k'] = k.numpy()
        out[f'upfirdn2d/{name}/args'] = np.array([up, up, down, down, *pad], dtype=np.int64)
        out[f'upfirdn2d/{name}/y'] = ref.contiguous().numpy()

    # fused_leaky_relu: restated from the kernel (no reference CPU implementation exists)
    x = torch.randn(3, 5, 4, 6, generator=g)
    b = torch.randn(5, generator=g)
    out['fused/x'] = x.numpy()
    out['fused/b'] = b.numpy()
    out['fused/y'] = so.fused_leaky_relu(x, b).numpy()
    assert torch.equal(so.fused_leaky_relu(x, b), F.leaky_relu(x + b.view(1, -1, 1, 1), 0.2) * 2 ** 0.5)

    # ------------------------------------------------------------- generator
    for tag, size, sdim, n_mlp, batch, trunc, mix in (
            ('g32', 32, 512, 8, 2, 1.0, False),
            ('g16trunc', 16, 64, 2, 3, 0.7, False),
            ('g16mix', 16, 64, 2, 2, 0.7, True)):
        torch.manual_seed(0)
        ref_g = model.Generator(size, sdim, n_mlp, channel_multiplier=2).eval()
        spec = so.GeneratorSpec(size, sdim, n_mlp, 2)
        sd = so.init_state_dict(spec, seed=0)
        ref_sd = ref_g.state_dict()
        assert set(ref_sd.keys()) == set(sd.keys()), set(ref_sd.keys()) ^ set(sd.keys())
        for k_ in ref_sd:
            eq(sd[k_], ref_sd[k_], f'init {tag} {k_}')
        so.perturb_zero_params(sd, seed=1234)
        ref_g.load_state_dict(sd)

        torch.manual_seed(1)  # build_latent_and_noise_generator semantics, utils/dataset_creation.py:32-37
        z = torch.randn(batch, sdim)
        noise = ref_g.make_noise()
        torch.manual_seed(1)
        z2 = torch.randn(batch, sdim)
        noise2 = so.make_noise(spec)
        eq(z, z2, 'latent stream')
        for a, b_ in zip(noise, noise2):
            eq(a, b_, 'noise stream')
        styles = [z]
        kwargs = dict(noise=noise, return_intermediate_activations=True)
        if trunc < 1:
            torch.manual_seed(7)
            ml_ref = ref_g.mean_latent(64)
            torch.manual_seed(7)
            ml = so.mean_latent(sd, spec, 64)
            eq(ml, ml_ref, 'mean_latent')
            kwargs.update(truncation=trunc, truncation_latent=ml)
            out[f'{tag}/mean_latent'] = ml.numpy()
        if mix:
            zb = torch.randn(batch, sdim)
            styles = [z, zb]
            kwargs.update(inject_index=3)
            out[f'{tag}/z_b'] = zb.numpy()
        with torch.no_grad():
            img_ref, acts_ref = ref_g(styles, **kwargs)
        img, acts = so.generator_forward(sd, spec, styles, **kwargs)
        eq(img, img_ref, f'{tag} image')
        assert sorted(acts) == sorted(acts_ref)
        for k_ in acts_ref:
            eq(acts[k_], acts_ref[k_], f'{tag} act {k_}')
        # return_latents path and the randomize_noise=False buffers
        with torch.no_grad():
            i2_ref, lat_ref = ref_g(styles, **{**kwargs, 'return_intermediate_activations': False,
                                              'return_latents': True, 'noise': None, 'randomize_noise': False})
        i2, lat = so.generator_forward(sd, spec, styles, **{**kwargs, 'return_intermediate_activations': False,
                                                            'return_latents': True, 'noise': None,
                                                            'randomize_noise': False})
        eq(i2, i2_ref, f'{tag} image (buffer noise)')
        eq(lat, lat_ref, f'{tag} latent')

        out[f'{tag}/config'] = np.array([size, sdim, n_mlp, 2, batch], dtype=np.int64)
        out[f'{tag}/truncation'] = np.array([trunc], dtype=np.float64)
        out[f'{tag}/z'] = z.numpy()
        out[f'{tag}/image'] = img_ref.numpy()
        out[f'{tag}/image_buffer_noise'] = i2_ref.numpy()
        out[f'{tag}/latent'] = lat_ref.numpy()
        for k_, a in acts_ref.items():
            flat = a.reshape(-1).numpy()
            idx = sample_idx(flat.size, 2048, seed=k_)
            out[f'{tag}/act{k_}/shape'] = np.array(a.shape, dtype=np.int64)
            out[f'{tag}/act{k_}/idx'] = idx
            out[f'{tag}/act{k_}/val'] = flat[idx]
            out[f'{tag}/act{k_}/sum_abs'] = np.array([np.abs(flat.astype(np.float64)).sum()])
        # a couple of weight samples so the tests can prove the seeded init is the reference's
        for k_ in ('style.1.weight', 'conv1.conv.weight', f'convs.{2 * (spec.log_size - 3)}.conv.weight',
                   'noises.noise_2'):
            flat = sd[k_].reshape(-1).numpy()
            idx = sample_idx(flat.size, 64, seed=3)
            out[f'{tag}/w/{k_}/idx'] = idx
            out[f'{tag}/w/{k_}/val'] = flat[idx]

        if tag == 'g32':
            g32 = (sd, spec, acts_ref)

    # -------------------------------------------------------------- labelling
    fc, base, bw = import_reference_labelling()
    sd, spec, acts = g32
    torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU here; factor_catalog.py:61 calls .cuda()
    layers = {'4': 4, '5': 7, '6': 4, '7': 20}   # layer -> k  (16^2 and 32^2 maps of the size-32 generator)
    gk = torch.Generator().manual_seed(5)
    catalog_ref, catalog = {}, {}
    for layer, k in layers.items():
        a = acts[int(layer)]
        flat = lo.partial_flat(a)
        pick = torch.randperm(flat.shape[0], generator=gk)[:k]
        cent = F.normalize(flat[pick] + 0.05 * torch.randn(k, flat.shape[1], generator=gk), dim=1)
        f = fc.FactorCatalog(k)
        f._factorization.cluster_centers_ = cent.numpy()
        catalog_ref[layer] = f
        catalog[layer] = cent
        ids_ref = f.predict(a)
        ids, margin = lo.predict_with_margin(a, cent)
        eq(ids, ids_ref, f'predict layer {layer}')
        eq(lo.predict(a, cent), ids_ref, f'predict layer {layer}')
        out[f'label/{layer}/centroids'] = cent.numpy()
        out[f'label/{layer}/ids'] = ids_ref.numpy().astype(np.uint8)
        out[f'label/{layer}/margin'] = margin.numpy()
    names = ['background', 'printed_text', 'handwritten_text']
    raw_map = {layer: {str(c): names[(c * 7 + int(layer)) % 3] for c in range(k)} for layer, k in layers.items()}
    inv = lo.invert_class_label_map(raw_map)

    seg = bw.BlackWhiteHandwrittenPrintedTextDatasetSegmenter.__new__(
        bw.BlackWhiteHandwrittenPrintedTextDatasetSegmenter)
    seg.catalog = catalog_ref
    seg.image_size = 32
    seg.debug = False
    seg.class_to_color_map = {n: (0, 0, 0) for n in names}
    seg.keys_to_merge = {'merged': ['4', '6']}
    # reference's own inversion (base_cluster_based_dataset_segmenter.py:56-67) run on a temp file
    import json
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        seg.base_dir = Path(td)
        seg.num_clusters = 4
        (Path(td) / 'merged_classes_4.json').write_text(json.dumps(raw_map))
        inv_ref = seg.load_class_label_map()
    assert {k: dict(v) for k, v in inv_ref.items()} == {k: dict(v) for k, v in inv.items()}
    pc_ref = seg.prepare_image_segmentation(acts, inv_ref)
    pc_ref = seg.merge_sub_images(pc_ref)
    pc = lo.prepare_image_segmentation(acts, catalog, inv, 32)
    pc = lo.merge_sub_images(pc, {'merged': ['4', '6']}, names)
    assert set(pc) == set(pc_ref)
    for layer in pc_ref:
        for cn in pc_ref[layer]:
            eq(pc[layer][cn], pc_ref[layer][cn], f'mask {layer}/{cn}')
            out[f'label/mask/{layer}/{cn}'] = np.packbits(pc_ref[layer][cn].numpy())
    out['label/class_map_json'] = np.frombuffer(json.dumps(raw_map).encode(), dtype=np.uint8)

    # bilinear-then-assign mode: reference op = torch.nn.Upsample(scale_factor, 'bilinear') (create_dataset...:43)
    a = acts[4]
    up = torch.nn.Upsample(scale_factor=32 / a.shape[-1], mode='bilinear')
    ids_ref = catalog_ref['4'].predict(up(a))
    eq(lo.bilinear_then_predict(a, catalog['4'], 32), ids_ref, 'bilinear_then_predict')
    out['label/4/ids_bilinear'] = ids_ref.numpy().astype(np.uint8)

    np.savez_compressed(HERE / 'golden_v1.npz', **out)
    size = (HERE / 'golden_v1.npz').stat().st_size
    print(f'wrote {HERE / "golden_v1.npz"}: {len(out)} arrays, {size / 1024:.1f} KiB')


if __name__ == '__main__':
    main()
