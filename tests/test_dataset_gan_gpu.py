"""GPU parity of the DatasetGAN labeller through the C-ABI: against golden results of the
reference's own classes (tests/golden/make_golden_dataset_gan.py) and against the oracle at 256^2.
Tolerance: labels / votes equal wherever the smallest top-2 logit margin over the networks exceeds 1e-3, and >= 99.9 %
of all pixels (the first Linear runs as a 3-term bf16 split at native resolution before the upsample; see
csrc/dataset_gan.cu)."""
import os

import numpy
import pytest
import torch

from oracle import dataset_gan_oracle as dg
from oracle import stylegan2_oracle as so
from synthesis_in_style_b200 import dataset_gan as pg
from synthesis_in_style_b200.model import Generator

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
COLORS = {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'}


def build_ensemble(states, n_class, feature_size):
    ens = pg.PixelEnsembleClassifier(n_class, 0, 0)
    for st in states:
        net = pg.PixelClassifier(n_class, feature_size)
        net.load_state_dict(st)
        net.eval()
        ens.add_network(net)
    return ens


def agreement(got, want, margin, tol=1e-3):
    safe = margin > tol
    return bool((got[safe] == want[safe]).all()), float((got == want).float().mean()), int(safe.sum())


def test_labels_equal_reference_golden(cuda_device):
    gold = numpy.load(os.path.join(HERE, 'golden', 'golden_dataset_gan_v1.npz'))
    size, batch, n_class, feature_size = (int(v) for v in gold['cfg'])
    spec = so.GeneratorSpec(size, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    z = torch.from_numpy(gold['z'])
    noise = [torch.from_numpy(gold[f'noise/{i}']) for i in range(spec.num_layers)]
    _, acts = so.generator_forward(sd, spec, [z], noise=noise, return_intermediate_activations=True)
    states = [dg.init_classifier_state(feature_size, n_class, seed=40 + i, base_seed=39) for i in range(3)]
    checksum = numpy.array([float(sum(v.double().sum() for v in st.values())) for st in states])
    assert numpy.allclose(checksum, gold['classifier_checksum'], rtol=0, atol=1e-6), 'classifier seeds drifted'
    ens = build_ensemble(states, n_class, feature_size)
    dev_acts = {k: v.to(cuda_device) for k, v in acts.items()}
    seg = pg.DatasetGANSegmenter(None, size, COLORS, ensemble=ens)
    labels, votes, colors = ens.predict_label_images(dev_acts, size, colors=list(seg.class_to_color_map.values()), want_votes=True)
    ens.check(cuda_device)
    margin = torch.from_numpy(gold['min_margin'])
    ok, frac, n_safe = agreement(labels.cpu(), torch.from_numpy(gold['labels']), margin)
    assert ok and frac >= 0.999, (ok, frac)
    assert n_safe > 0.9 * margin.numel()
    want_votes = torch.from_numpy(gold['votes'])
    for m in range(3):
        okv, fracv, _ = agreement(votes.cpu()[..., m], want_votes[..., m], margin)
        assert okv and fracv >= 0.999, (m, fracv)
    # colour images follow the labels exactly (dataset_gan_segmenter.py:43-53)
    same = labels.cpu() == torch.from_numpy(gold['labels'])
    assert (colors.cpu().numpy()[same.numpy()] == gold['color_images'][same.numpy()]).all()
    images, drop = seg.create_segmentation_image(dev_acts)
    assert drop == [] and images.dtype == numpy.uint8 and images.shape == (batch, size, size, 3)
    assert (images == colors.cpu().numpy()).all()
    assert (seg.label_images_to_color_images(labels) == images).all()


def test_256_against_oracle(cuda_device):
    """BASELINE shape (256^2, 14 captures, F = 5888, 3 networks) on one sample: captures from the B200 generator, labels vs
    the oracle's materialised-feature path on the same captures."""
    size = 256
    spec = so.GeneratorSpec(size, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    g = Generator(size, 512, 8)
    g.load_state_dict(sd)
    g = g.to(cuda_device).eval()
    torch.manual_seed(3)
    with torch.no_grad():
        _, acts = g([torch.randn(1, 512).to(cuda_device)], return_intermediate_activations=True, noise=[n.to(cuda_device) for n in so.make_noise(spec)])
    feature_size = sum(a.shape[1] for a in acts.values())
    assert feature_size == 5888 and pg.get_dataset_gan_params(acts, size)['feature_size'] == 5888
    states = [dg.init_classifier_state(feature_size, 3, seed=50 + i, base_seed=49) for i in range(3)]
    ens = build_ensemble(states, 3, feature_size)
    labels, votes, _ = ens.predict_label_images(acts, size, want_votes=True)
    ens.check(cuda_device)
    cpu_acts = {k: v.cpu() for k, v in acts.items()}
    want, margin, want_votes = dg.predict_labels([dg.ClassifierParams(s) for s in states], cpu_acts, size)
    ok, frac, n_safe = agreement(labels.cpu().float(), want, margin)
    assert ok and frac >= 0.999, (ok, frac)
    assert len(torch.unique(want)) >= 2
    for m in range(3):
        okv, fracv, _ = agreement(votes.cpu()[..., m].float(), want_votes[..., m], margin)
        assert okv and fracv >= 0.999, (m, fracv)


def test_checkpoint_loading_and_errors(cuda_device, tmp_path):
    feature_size, n_class = 64, 3
    states = [dg.init_classifier_state(feature_size, n_class, seed=i) for i in range(2)]
    torch.save({'network_0': states[0], 'network_1': states[1], 'optimizer_network_0': {}, 'iteration': 5}, tmp_path / 'ckpt.pt')
    seg = pg.DatasetGANSegmenter(None, 16, COLORS, classifier_path=str(tmp_path / 'ckpt.pt'), feature_size=feature_size)
    assert len(seg.ensemble.networks) == 2
    acts = {0: torch.randn(2, 32, 4, 4, device=cuda_device), 1: torch.randn(2, 32, 16, 16, device=cuda_device)}
    images, drop = seg.create_segmentation_image(acts)
    assert images.shape == (2, 16, 16, 3) and drop == []
    want, _, _ = dg.predict_labels([dg.ClassifierParams(s) for s in states], {k: v.cpu() for k, v in acts.items()}, 16)
    assert float((seg.predict_labels(acts).cpu().float() == want).float().mean()) >= 0.99
    # an image size that is not a multiple of the 16 x 8 tile takes the generic (untiled) tail kernel
    small = {0: torch.randn(2, 32, 4, 4, device=cuda_device), 1: torch.randn(2, 32, 8, 8, device=cuda_device)}
    got8, votes8, _ = seg.ensemble.predict_label_images(small, 8, want_votes=True)
    want8, margin8, want_votes8 = dg.predict_labels([dg.ClassifierParams(s) for s in states], {k: v.cpu() for k, v in small.items()}, 8)
    safe8 = margin8 > 1e-3
    assert bool((got8.cpu().float()[safe8] == want8[safe8]).all()) and votes8.shape == (2, 8, 8, 2)
    with pytest.raises(RuntimeError):                      # feature size mismatch
        seg.ensemble.predict_label_images({0: torch.randn(2, 32, 4, 4, device=cuda_device)}, 16)
    with pytest.raises(RuntimeError):                      # host tensors are refused: no CPU fallback
        seg.ensemble.predict_label_images({0: torch.randn(2, 64, 4, 4)}, 16)


def test_weight_updates_between_predicts_are_seen(cuda_device):
    """The native ensemble is rebuilt when a network's tensors change in place or are replaced (load_state_dict, copy_):
    labels after an update equal those of a fresh ensemble with the new weights, not the stale ones."""
    size, n_class = 32, 3
    spec = so.GeneratorSpec(size, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    g = Generator(size, 512, 8)
    g.load_state_dict(sd)
    g = g.to(cuda_device).eval()
    torch.manual_seed(3)
    with torch.no_grad():
        _, acts = g([torch.randn(2, 512, device=cuda_device)], return_intermediate_activations=True)
    feat = sum(t.shape[1] for t in acts.values())
    old_states = [dg.init_classifier_state(feat, n_class, seed=60 + i, base_seed=59) for i in range(2)]
    new_states = [dg.init_classifier_state(feat, n_class, seed=70 + i, base_seed=69) for i in range(2)]
    ens = build_ensemble(old_states, n_class, feat)
    before = ens.predict_label_images(acts, size)[0].clone()
    nets = list(ens.get_networks().values())
    nets[0].load_state_dict(new_states[0])                       # in-place copy into the existing tensors
    with torch.no_grad():
        for k, v in new_states[1].items():
            nets[1].state_dict()[k].copy_(v)                     # raw in-place update
    after = ens.predict_label_images(acts, size)[0]
    fresh = build_ensemble(new_states, n_class, feat).predict_label_images(acts, size)[0]
    assert torch.equal(after, fresh)
    assert not torch.equal(after, before)
