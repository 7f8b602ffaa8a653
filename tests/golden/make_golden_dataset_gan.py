#!/usr/bin/env python
"""Generate tests/golden/golden_dataset_gan_v1.npz by running the REFERENCE's DatasetGAN labeller classes in this
container (needs /root/reference).  The reference hard-codes CUDA placement (`.cuda()`, `device='cuda'`); there is no
GPU here, so those placements are neutralised for the duration of the run (Tensor.cuda / Module.cuda -> identity,
device='cuda' -> 'cpu') -- the arithmetic that runs is the reference's, in fp32 on the CPU.

Stored: the latent / noise of a 32^2 generator run (captures are recomputed by the oracle generator), checksums of three
synthetic PixelClassifier state dicts (re-drawn from their seeds), and the reference's outputs: scale_activations features checksum, per-network votes, label images, colour
images.  Asserts that oracle/dataset_gan_oracle.py reproduces them exactly.
Usage: python tests/golden/make_golden_dataset_gan.py
"""
import importlib
import os
import sys
import types

import numpy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/stylegan_code_finder'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import dataset_gan_oracle as dg  # noqa: E402
from oracle import stylegan2_oracle as so  # noqa: E402

COLORS = {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'}


def import_reference():
    """networks/__init__.py and data/__init__.py pull in packages this image lacks; register bare packages so that only
    the three modules on this path are imported."""
    for name, sub in (('networks', 'networks'), ('networks.pixel_classifier', 'networks/pixel_classifier'), ('data', 'data')):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, sub)]
            sys.modules[name] = m
    # data.dataset_gan_dataset imports its dataset base class (h5py, ...): give it a stub, only scale_activations is used
    stub = types.ModuleType('data.base_dataset_gan_dataset')
    stub.BaseDatasetGANDataset = object
    sys.modules['data.base_dataset_gan_dataset'] = stub
    model = importlib.import_module('networks.pixel_classifier.model')
    ds = importlib.import_module('data.dataset_gan_dataset')
    seg = importlib.import_module('segmentation.dataset_gan_segmenter')
    return model, ds, seg


class cpu_placement:
    def __enter__(self):
        self.saved = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.zeros, torch.empty)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        def on_cpu(fn):
            def wrapped(*a, **k):
                if k.get('device') == 'cuda':
                    k['device'] = 'cpu'
                return fn(*a, **k)
            return wrapped
        torch.zeros, torch.empty = on_cpu(self.saved[2]), on_cpu(self.saved[3])

    def __exit__(self, *exc):
        torch.Tensor.cuda, torch.nn.Module.cuda, torch.zeros, torch.empty = self.saved


def main():
    model, ds, seg_mod = import_reference()
    size, batch, n_class = 32, 3, 3
    spec = so.GeneratorSpec(size, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    torch.manual_seed(7)
    z = torch.randn(batch, 512)
    noise = so.make_noise(spec)
    _, acts = so.generator_forward(sd, spec, [z], noise=noise, return_intermediate_activations=True)
    feature_size = sum(a.shape[1] for a in acts.values())
    states = [dg.init_classifier_state(feature_size, n_class, seed=40 + i, base_seed=39) for i in range(3)]

    out = {'cfg': numpy.array([size, batch, n_class, feature_size], dtype=numpy.int64), 'z': z.numpy()}
    for i, n in enumerate(noise):
        out[f'noise/{i}'] = n.numpy()
    # the classifier weights are re-drawn by the tests (dg.init_classifier_state, same seeds); keep a checksum only
    out['classifier_checksum'] = numpy.array([float(sum(v.double().sum() for v in st.values())) for st in states])

    with cpu_placement():
        upsamplers = [torch.nn.Upsample(scale_factor=size / a.shape[-1], mode='bilinear') for a in acts.values()]
        feats = ds.scale_activations([acts], upsamplers)[0]
        segmenter = seg_mod.DatasetGANSegmenter.__new__(seg_mod.DatasetGANSegmenter)
        segmenter.image_size = size
        segmenter.class_to_color_map = segmenter.load_class_to_color_map(COLORS)
        segmenter.upsamplers = upsamplers
        ensemble = model.PixelEnsembleClassifier(n_class, size, 0)
        for st in states:
            net = model.PixelClassifier(n_class, feature_size)
            net.load_state_dict(st)
            net.eval()
            ensemble.add_network(net)
        segmenter.ensemble = ensemble
        with torch.no_grad():
            votes = torch.stack([net.predict_classes(feats.reshape(-1, feature_size)).squeeze() for net in ensemble.networks.values()], dim=1)
            color_images, drop = segmenter.create_segmentation_image(acts)
            labels = segmenter.predict_labels(feats)
    assert drop == []

    mine_feats = dg.scale_activations(acts, size)
    assert torch.equal(mine_feats, feats)
    models = [dg.ClassifierParams(st) for st in states]
    mine_labels, margins, mine_votes = dg.predict_labels(models, acts, size)
    assert torch.equal(mine_votes.reshape(-1, 3), votes.float()), int((mine_votes.reshape(-1, 3) != votes.float()).sum())
    assert torch.equal(mine_labels, labels)
    mine_color, _ = dg.create_segmentation_image(models, acts, size, segmenter.class_to_color_map)
    assert numpy.array_equal(mine_color, color_images)

    out['votes'] = votes.numpy().astype(numpy.uint8).reshape(batch, size, size, 3)
    out['labels'] = labels.numpy().astype(numpy.uint8)
    out['color_images'] = color_images
    out['min_margin'] = margins.numpy().astype(numpy.float32)
    out['features_sum'] = feats.double().sum(dim=(0, 1, 2)).numpy()
    path = os.path.join(HERE, 'golden_dataset_gan_v1.npz')
    numpy.savez_compressed(path, **out)
    hist = numpy.bincount(out['labels'].ravel(), minlength=n_class)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB; label histogram', hist.tolist(),
          'vote disagreement', float((votes.min(1).values != votes.max(1).values).float().mean()))


if __name__ == '__main__':
    main()
