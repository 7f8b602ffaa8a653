"""Contour stage on the device (csrc/contours.cu through sis_contour_stage) against
  * the golden label images + drop lists produced by the reference's own segmenter classes
    (tests/golden/golden_contours_v1.npz, made by tests/golden/make_golden_contours.py), byte for byte, and
  * the polygon implementation (synthesis_in_style_b200/contours.py, itself pinned by those goldens) on document-like
    and noise-like masks, empty / full masks, single-key and three-key configurations.
The device's own output (before any host fallback) is checked: label images must be equal for EVERY image, flags must be
0 / 1 as the reference decides, or 2 (undecided: the drop rule depends on contour order) -- and 2 must stay rare."""
import os

import numpy
import pytest
import torch

from oracle import contour_oracle as co
from synthesis_in_style_b200 import contours as pc
from synthesis_in_style_b200 import contours_device as pd

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
COLORS = {'background': (0, 0, 0), 'printed_text': (0, 0, 255), 'handwritten_text': (255, 0, 0)}


@pytest.fixture(scope='module')
def gold():
    return numpy.load(os.path.join(HERE, 'golden', 'golden_contours_v1.npz'))


def to_stacked(pred, device):
    return {key: (list(per_class), torch.stack([torch.from_numpy(numpy.ascontiguousarray(m).astype(numpy.uint8)) for m in per_class.values()]).to(device))
            for key, per_class in pred.items()}


def check_against(pred, batch, cfg, device, want_images, want_drop, max_undecided=None, allow_unsettled=False):
    stage = pd.DeviceContourStage(cfg)
    images_d, flags_d = stage.run(to_stacked(pred, device))
    torch.cuda.synchronize()
    images, flags = images_d.cpu().numpy(), flags_d.cpu().numpy()
    reason = stage.last_info[2]
    assert reason in ((0, 3) if allow_unsettled else (0,)), f'device stage gave the batch up (reason {reason})'
    assert images.shape == want_images.shape and images.dtype == numpy.uint8
    if reason == 3:        # the merge fixpoint was still moving after the enqueued rounds: every image goes to the host path
        assert (flags == pd.FLAG_HOST).all()
    else:
        bad = [b for b in range(batch) if not numpy.array_equal(images[b], want_images[b])]
        assert not bad, f'label images differ for images {bad}'
        for b in range(batch):
            if flags[b] != pd.FLAG_HOST:
                assert (flags[b] == pd.FLAG_DROP) == (b in want_drop), f'image {b}: flag {flags[b]}, reference drop {b in want_drop}'
    undecided = int((flags == pd.FLAG_HOST).sum())
    if max_undecided is not None:
        assert undecided <= max_undecided, f'{undecided} of {batch} images undecided'
    # the public entry point resolves the undecided ones on the host
    got_images, got_drop = pd.segment(to_stacked(pred, device), batch, cfg)
    assert numpy.array_equal(got_images, want_images) and got_drop == sorted(want_drop)
    return undecided


def load_full(gold, tag):
    seed, batch, size, keep, min_area = (int(v) for v in gold[f'full/{tag}/cfg'])
    pred = {}
    for key in ('8', '9', '12', '13'):
        pred[key] = {name: numpy.unpackbits(gold[f'full/{tag}/mask/{key}/{name}'], axis=-1)[..., :size] for name in COLORS}
    return pred, batch, pc.ContourConfig(size, COLORS, ['8', '9'], ['12', '13'], bool(keep), min_area)


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd', 'e'])
def test_device_equals_reference_goldens(gold, cuda_device, tag):
    pred, batch, cfg = load_full(gold, tag)
    check_against(pred, batch, cfg, cuda_device, gold[f'full/{tag}/images'], list(gold[f'full/{tag}/drop']))


def noise_masks(seed, batch, size, smooth=(3, 1)):
    """argmax of smoothed noise fields: dense specks, chains of overlaps, holes inside unions."""
    from scipy import ndimage
    rng = numpy.random.default_rng(seed)
    out = {}
    for key in ('8', '9', '12', '13'):
        s = smooth[0] if key in ('8', '9') else smooth[1]
        ids = ndimage.gaussian_filter(rng.standard_normal((batch, 3, size, size)), (0, 0, s, s)).argmax(1)
        out[key] = {name: (ids == j).astype(numpy.uint8) for j, name in enumerate(COLORS)}
    return out


@pytest.mark.parametrize('keep', [True, False])
@pytest.mark.parametrize('min_area', [0, 50])
def test_device_equals_polygon_path_documents(cuda_device, keep, min_area):
    batch, size = 8, 256
    pred = co.synthetic_document_masks(31 + min_area, batch, size)
    cfg = pc.ContourConfig(size, COLORS, ['8', '9'], ['12', '13'], keep, min_area)
    want_images, want_drop = pc.segment_masks(pred, batch, cfg)
    check_against(pred, batch, cfg, cuda_device, want_images, want_drop)


@pytest.mark.parametrize('case', [(5, 128, (3, 1), True, 10), (6, 96, (2, 1), False, 0), (7, 128, (4, 2), False, 50), (8, 64, (1, 1), True, 0)])
def test_device_equals_polygon_path_noise(cuda_device, case):
    seed, size, smooth, keep, min_area = case
    batch = 6
    pred = noise_masks(seed, batch, size, smooth)
    cfg = pc.ContourConfig(size, COLORS, ['8', '9'], ['12', '13'], keep, min_area)
    want_images, want_drop = pc.segment_masks(pred, batch, cfg)
    check_against(pred, batch, cfg, cuda_device, want_images, want_drop, allow_unsettled=True)


def test_device_edge_cases(cuda_device):
    size, batch = 64, 4
    zeros, ones = numpy.zeros((batch, size, size), numpy.uint8), numpy.ones((batch, size, size), numpy.uint8)
    ring = zeros.copy()
    ring[:, 8:56, 8:56] = 1
    ring[:, 10:54, 10:54] = 0                       # a ring: its hole is part of the filled contour
    dot = zeros.copy()
    dot[:, 30:33, 30:33] = 1                        # inside the ring's hole
    cases = {
        'empty': {k: {'background': ones, 'printed_text': zeros, 'handwritten_text': zeros} for k in ('8', '9', '12', '13')},
        'full': {k: {'background': zeros, 'printed_text': ones, 'handwritten_text': zeros} for k in ('8', '9', '12', '13')},
        'ring': {'8': {'background': zeros, 'printed_text': ring, 'handwritten_text': zeros},
                 '9': {'background': zeros, 'printed_text': dot, 'handwritten_text': ring},
                 '12': {'background': zeros, 'printed_text': ring, 'handwritten_text': zeros},
                 '13': {'background': zeros, 'printed_text': dot, 'handwritten_text': zeros}},
    }
    for name, pred in cases.items():
        for keep in (True, False):
            cfg = pc.ContourConfig(size, COLORS, ['8', '9'], ['12', '13'], keep, 0)
            want_images, want_drop = pc.segment_masks(pred, batch, cfg)
            check_against(pred, batch, cfg, cuda_device, want_images, want_drop, max_undecided=0)


@pytest.mark.parametrize('keys', [(['8'], ['12']), (['8', '9', '12'], ['12', '13', '9'])])
def test_device_other_key_counts(cuda_device, keys):
    batch, size = 4, 128
    pred = co.synthetic_document_masks(77, batch, size)
    cfg = pc.ContourConfig(size, COLORS, keys[0], keys[1], False, 10)
    want_images, want_drop = pc.segment_masks(pred, batch, cfg)
    check_against(pred, batch, cfg, cuda_device, want_images, want_drop)


@pytest.mark.parametrize('size', [512, 1024])
def test_device_large_images(cuda_device, size):
    """512^2 (bitmasks of the widest windows still in shared memory) and 1024^2 (workspace slices, windows without rim
    columns): a frame around the whole image makes one group span it; a blob inside the frame's hole gets swallowed."""
    batch = 2
    pred = co.synthetic_document_masks(5, batch, size)
    pred = {k: {n: numpy.ascontiguousarray(m).astype(numpy.uint8) for n, m in v.items()} for k, v in pred.items()}
    frame = numpy.zeros((size, size), numpy.uint8)
    frame[0:6, :] = frame[-6:, :] = 1
    frame[:, 0:6] = frame[:, -6:] = 1
    for k in ('12', '13', '8', '9'):
        pred[k]['printed_text'][1] |= frame
        pred[k]['background'][1] &= 1 - frame
    cfg = pc.ContourConfig(size, COLORS, ['8', '9'], ['12', '13'], False, 50)
    want_images, want_drop = pc.segment_masks(pred, batch, cfg)
    check_against(pred, batch, cfg, cuda_device, want_images, want_drop)


def test_device_stage_errors(cuda_device):
    cfg = pc.ContourConfig(64, COLORS, ['8', '9'], ['12', '13'], True, 0)
    pred = co.synthetic_document_masks(3, 2, 64)
    stage = pd.DeviceContourStage(cfg)
    with pytest.raises(RuntimeError):
        stage.run({k: (list(v), torch.stack([torch.from_numpy(m.astype(numpy.uint8)) for m in v.values()])) for k, v in pred.items()})   # host tensors
    with pytest.raises(KeyError):
        pd.DeviceContourStage(cfg, fine_class='no_such_class')


def test_device_randomised_cross_check(cuda_device):
    """A short run of scripts/stress_contours.py (random mask statistics, sizes that are not multiples of 32, 1-3 keys per
    stage, all configurations): every case must agree with the host polygon path.  (530 cases on a B200:
    profiles/r03t_contour_stress.txt.)"""
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    res = subprocess.run([sys.executable, os.path.join(root, 'scripts', 'stress_contours.py'), '--cases', '30', '--seed', '11'],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert ' 0 mismatching cases' in res.stdout
