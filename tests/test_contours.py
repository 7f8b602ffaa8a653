"""Contour post-processing: the product implementation against golden results produced by the
reference's own classes (tests/golden/make_golden_contours.py), against the reference's merge fixtures, and against the
oracle restatement on random cases.  CPU only."""
import os
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy
import pytest

from oracle import contour_oracle as co
from synthesis_in_style_b200 import contours as pc

HERE = os.path.dirname(os.path.abspath(__file__))
COLORS = {'background': (0, 0, 0), 'printed_text': (0, 0, 255), 'handwritten_text': (255, 0, 0)}


@pytest.fixture(scope='module')
def gold():
    return numpy.load(os.path.join(HERE, 'golden', 'golden_contours_v1.npz'))


def contours_of(filled_masks):
    return [cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)[0][0] for m in filled_masks]


def filled(contours, size):
    out = numpy.zeros((len(contours), size, size), dtype=numpy.uint8)
    for i, c in enumerate(contours):
        cv2.drawContours(out[i], [c], 0, 1, cv2.FILLED)
    return out


def canonical(masks):
    keys = [(int(m.sum()), int(numpy.flatnonzero(m)[0]) if m.any() else -1) for m in masks]
    order = sorted(range(len(masks)), key=lambda i: keys[i])
    return masks[order] if len(order) else masks


def merge_inputs(gold, name):
    n_sub = int(gold[f'merge/{name}/n_sub'])
    return {str(i): {'printed_text': [contours_of(gold[f'merge/{name}/in{i}'])]} for i in range(n_sub)}


@pytest.mark.parametrize('name', ['two', 'three', 'one_empty', 'all_empty', 'no_overlap2', 'no_overlap3'])
@pytest.mark.parametrize('keep', [True, False])
@pytest.mark.parametrize('impl', ['product', 'oracle'])
def test_reference_merge_fixtures(gold, name, keep, impl):
    """The polygon cases of the reference's tests/test_merge_contours.py, results as the reference computed them."""
    per_sub = merge_inputs(gold, name)
    fn = pc.merge_contours_of_same_class_from_different_images if impl == 'product' else co.merge_across_sub_images
    got = fn(per_sub, 1, keep, ('printed_text',))['printed_text'][0]
    tag = f'merge/{name}/keep{int(keep)}'
    if tag + '/none' in gold.files:
        assert got is None
        return
    want = gold[tag]
    assert got is not None and len(got) == len(want)
    assert numpy.array_equal(canonical(filled(list(got), 1024)), want)


@pytest.mark.parametrize('name', ['two', 'three'])
def test_reference_merge_same_image(gold, name):
    per_sub = merge_inputs(gold, name)
    flat = {'printed_text': [[c for sub in per_sub.values() for c in sub['printed_text'][0]]]}
    got = pc.merge_contours_of_same_class_from_same_image(flat)['printed_text'][0]
    assert numpy.array_equal(canonical(filled(got, 1024)), gold[f'merge/{name}/same_image'])


def test_overlap_primitives():
    """test_merge_contours.py TestOverlapDetection: triangles that do / do not overlap."""
    def tri(points):
        img = numpy.zeros((64, 64), dtype=numpy.uint8)
        cv2.fillPoly(img, [numpy.array(points, dtype=numpy.int32)], 255)
        return cv2.findContours(img, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)[0][0]
    a, b = tri([(5, 5), (15, 5), (5, 15)]), tri([(5, 30), (30, 5), (30, 30)])
    assert pc.merge_two_contours_if_overlapping(a, b) is None and pc.contour_overlap(a, b) == 0
    c, d = tri([(40, 40), (5, 40), (40, 5)]), tri([(30, 30), (25, 30), (30, 25)])
    merged = pc.merge_two_contours_if_overlapping(c, d)
    assert merged is not None and len(merged) == 1
    assert pc.contour_overlap(c, d) == co.contour_overlap(c, d) > 0


def load_full(gold, tag):
    seed, batch, size, keep, min_area = (int(v) for v in gold[f'full/{tag}/cfg'])
    pred = {}
    for key in ('8', '9', '12', '13'):
        pred[key] = {name: numpy.unpackbits(gold[f'full/{tag}/mask/{key}/{name}'], axis=-1)[..., :size].astype(bool) for name in COLORS}
    cfg = pc.ContourConfig(size, COLORS, ['8', '9'], ['12', '13'], bool(keep), min_area)
    return pred, batch, cfg


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd', 'e'])
def test_label_images_equal_reference(gold, tag):
    """Colour label images + drop lists exactly as the reference's create_segmentation_image produced them."""
    pred, batch, cfg = load_full(gold, tag)
    images, drop = pc.segment_masks(pred, batch, cfg)
    assert images.dtype == numpy.uint8 and images.shape == gold[f'full/{tag}/images'].shape
    assert numpy.array_equal(images, gold[f'full/{tag}/images'])
    assert sorted(drop) == list(gold[f'full/{tag}/drop'])


def test_parallel_equals_serial(gold):
    pred, batch, cfg = load_full(gold, 'e')
    with ThreadPoolExecutor(4) as pool:
        images, drop = pc.segment_masks_parallel(pred, batch, cfg, pool)
    assert numpy.array_equal(images, gold['full/e/images']) and sorted(drop) == list(gold['full/e/drop'])


def test_oracle_equals_reference_small(gold):
    """The slow restatement on the smallest golden case (the generator script asserts all of them)."""
    pred, batch, cfg = load_full(gold, 'd')
    images, drop = co.create_segmentation_image(pred, batch, cfg.image_size, COLORS, ['8', '9'], ['12', '13'],
                                                cfg.only_keep_overlapping, cfg.min_class_contour_area)
    assert numpy.array_equal(images, gold['full/d/images']) and sorted(drop) == list(gold['full/d/drop'])


def random_blobs(rng, n, size):
    out = []
    for _ in range(n):
        img = numpy.zeros((size, size), dtype=numpy.uint8)
        kind = rng.randint(3)
        x, y = rng.randint(0, size - 12, size=2)
        w, h = rng.randint(3, 28, size=2)
        if kind == 0:
            img[y:y + h, x:x + w] = 1
        elif kind == 1:
            cv2.ellipse(img, (int(x + w // 2), int(y + h // 2)), (int(w // 2 + 1), int(h // 2 + 1)), 0, 0, 360, 1, -1)
        else:                                                         # a ring: union fills can swallow what is inside
            cv2.rectangle(img, (int(x), int(y)), (int(min(size - 1, x + w + 8)), int(min(size - 1, y + h + 8))), 1, 2)
        found = cv2.findContours(img, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
        if found:
            out.append(found[0])
    return out


@pytest.mark.parametrize('seed', range(6))
@pytest.mark.parametrize('keep', [True, False])
def test_merge_order_and_shapes_equal_oracle(seed, keep):
    """merge_contours must reproduce the reference's merge ORDER too (the drop rule looks at the first contour)."""
    rng = numpy.random.RandomState(100 + seed)
    contours = random_blobs(rng, 18 + 4 * seed, 96)
    want = co.merge_contours(contours, keep)
    got = pc.merge_contours(contours, keep)
    assert len(got) == len(want)
    assert numpy.array_equal(filled(got, 140), filled(want, 140))      # same shapes in the same order


def test_empty_and_single_inputs():
    assert pc.merge_contours([]) == []
    one = random_blobs(numpy.random.RandomState(0), 1, 64)
    assert len(pc.merge_contours(one)) == 1 and pc.merge_contours(one, True) == []
    cfg = pc.ContourConfig(64, COLORS, ['8'], ['12'], True, 0)
    blank = {k: {n: numpy.zeros((2, 64, 64), dtype=bool) for n in COLORS} for k in ('8', '12')}
    images, drop = pc.segment_masks(blank, 2, cfg)
    assert images.shape == (2, 64, 64, 3) and not images.any() and drop == []


def test_device_stage_host_logic():
    """contours_device without a GPU: `supports` (every class under every key), the host fall-back entry point, and
    `segment` taking the host path as a whole when a key lacks a class (the device stage needs all of them)."""
    import torch
    from synthesis_in_style_b200 import contours_device as pd
    cfg = pc.ContourConfig(64, COLORS, ['8', '9'], ['12', '13'], False, 10)
    stage = pd.DeviceContourStage(cfg)
    names = {k: list(COLORS) for k in ('8', '9', '12', '13')}
    assert stage.classes == ['printed_text', 'handwritten_text'] and stage.supports(names)
    assert not stage.supports({**names, '9': ['background', 'printed_text']}) and not stage.supports({k: names[k] for k in ('8', '9', '12')})
    with pytest.raises(KeyError):
        pd.DeviceContourStage(cfg, fine_class='nope')
    pred = co.synthetic_document_masks(9, 3, 64)
    want_images, want_drop = pc.segment_masks(pred, 3, cfg)
    host = {k: (list(v), numpy.stack([m.astype(numpy.uint8) for m in v.values()])) for k, v in pred.items()}
    res = pd.host_fallback(host, [2, 0], cfg)
    assert set(res) == {0, 2}
    for b in (0, 2):
        assert numpy.array_equal(res[b][0], want_images[b]) and res[b][1] == (b in want_drop)
    # a key without 'handwritten_text': the reference's contour lists are then shorter; the whole batch goes to the host path
    partial = {k: {n: torch.from_numpy(m.astype(numpy.uint8)) for n, m in v.items() if not (k == '9' and n == 'handwritten_text')}
               for k, v in pred.items()}
    ref_images, ref_drop = pc.segment_masks({k: {n: m.numpy() for n, m in v.items()} for k, v in partial.items()}, 3, cfg)
    got_images, got_drop = pd.segment(partial, 3, cfg)
    assert numpy.array_equal(got_images, ref_images) and got_drop == sorted(ref_drop)
    assert pd.warm_worker() > 0


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd', 'e'])
def test_shape_algebra_equals_reference_goldens(gold, tag):
    """The algorithm the CUDA contour stage implements (oracle/contour_shape_oracle.py: filled shapes from labelled
    complements, area from crack / corner counts, merge as a fixpoint, per-pixel classification) reproduces the label
    images the reference's own classes produced, and its drop decision wherever it does not depend on contour order."""
    from oracle import contour_shape_oracle as so
    pred, batch, cfg = load_full(gold, tag)
    want_images, want_drop = gold[f'full/{tag}/images'], set(int(d) for d in gold[f'full/{tag}/drop'])
    undecided = 0
    for b in range(batch):
        masks = {k: {n: numpy.ascontiguousarray(m[b]).astype(numpy.uint8) for n, m in v.items()} for k, v in pred.items()}
        image, drop = so.segment_one(masks, cfg.keys_for_class_determination, cfg.keys_for_finegrained_segmentation, list(COLORS), COLORS,
                                     cfg.image_size, cfg.only_keep_overlapping, cfg.min_class_contour_area)
        assert numpy.array_equal(image, want_images[b]), (tag, b)
        if drop is None:
            undecided += 1
        else:
            assert drop == (b in want_drop), (tag, b)
    assert undecided <= batch // 2


def test_shape_algebra_identities_against_cv2():
    """The three identities on random masks: filled contour = complement component, contourArea = pixels - L/2 - 1,
    findContours order = descending first pixel."""
    from scipy import ndimage
    from oracle import contour_shape_oracle as so
    rng = numpy.random.RandomState(4)
    for trial in range(40):
        size = int(rng.choice([17, 32, 64]))
        if trial % 2:
            mask = (rng.rand(size, size) < rng.choice([0.05, 0.3, 0.6])).astype(numpy.uint8)
        else:
            mask = (ndimage.gaussian_filter(rng.randn(size, size), 2) > 0.02).astype(numpy.uint8)
        found, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        filled_map = so.fill_outside(mask > 0)
        labels, n = ndimage.label(filled_map, structure=numpy.ones((3, 3), int))
        assert n == len(found)
        firsts = []
        for contour in found:
            canvas = numpy.zeros_like(mask)
            cv2.drawContours(canvas, [contour], 0, 1, cv2.FILLED)
            ids = numpy.unique(labels[canvas > 0])
            assert len(ids) == 1 and numpy.array_equal(canvas > 0, labels == ids[0])
            count, chain = so.crack_stats(labels == ids[0])
            assert abs(cv2.contourArea(contour) - (count - chain / 2 - 1)) < 1e-9
            firsts.append(int(numpy.flatnonzero((labels == ids[0]).ravel())[0]))
        assert firsts == sorted(firsts, reverse=True)
