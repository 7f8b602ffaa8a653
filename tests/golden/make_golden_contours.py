#!/usr/bin/env python
"""Generate tests/golden/golden_contours_v1.npz by running the REFERENCE's own contour post-processing in this container
(needs /root/reference and cv2; the result travels as a small fixture, the reference does not).

  * merge fixtures: the polygon cases of the reference's tests/test_merge_contours.py, run through the reference's
    merge_contours_of_same_class_from_different_images / ..._from_same_image; stored as filled masks of the inputs and
    of the merged results (order-free comparison, as the reference's own `_results_equal` does);
  * full cases: synthetic document-like class masks for 2 class-determination + 2 fine-grained keys, run through the
    reference's BlackWhiteHandwrittenPrintedTextDatasetSegmenter.create_segmentation_image with
    prepare_image_segmentation replaced by "return these masks" -> colour label images + drop lists.
It also asserts that oracle/contour_oracle.py reproduces every stored result exactly.
Usage: python tests/golden/make_golden_contours.py
"""
import os
import sys
import types

import cv2
import numpy
import torch
from PIL import Image, ImageDraw

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/stylegan_code_finder'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import contour_oracle as co  # noqa: E402
from oracle.contour_oracle import synthetic_document_masks  # noqa: E402

COLORS = {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'}


def reference_segmenter(image_size, only_keep_overlapping, min_area, class_keys, fine_keys):
    from segmentation.black_white_handwritten_printed_text_segmenter import BlackWhiteHandwrittenPrintedTextDatasetSegmenter as Ref
    seg = Ref.__new__(Ref)                     # the constructor reads catalog pickles; the contour code needs none of it
    seg.image_size = image_size
    seg.debug = False
    seg.debug_images = {}
    seg.max_debug_text_size = 20
    seg.class_to_color_map = seg.load_class_to_color_map(COLORS)
    seg.class_id_map = seg.build_class_id_map(seg.class_to_color_map)
    seg.keys_for_class_determination = list(class_keys)
    seg.keys_for_finegrained_segmentation = list(fine_keys)
    seg.keys_to_merge = {}
    seg.only_keep_overlapping = only_keep_overlapping
    seg.min_class_contour_area = min_area
    seg.handwriting_overlap_threshold = 0.5
    seg.class_label_map = None
    return seg


def contour_from_polygon(polygon):
    x_max, y_max = numpy.asarray(polygon).max(axis=0)
    img = Image.new('L', (int(x_max) + 1, int(y_max) + 1))
    ImageDraw.Draw(img).polygon([tuple(p) for p in polygon], fill=255)
    return cv2.findContours(numpy.asarray(img), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)[0][0]


def filled(contours, size):
    out = numpy.zeros((len(contours), size, size), dtype=numpy.uint8)
    for i, c in enumerate(contours):
        cv2.drawContours(out[i], [c], 0, 1, cv2.FILLED)
    return out


def canonical(masks):
    """Order-free form of a set of filled masks: sorted by (area, first pixel index)."""
    keys = [(int(m.sum()), int(numpy.flatnonzero(m)[0]) if m.any() else -1) for m in masks]
    order = sorted(range(len(masks)), key=lambda i: keys[i])
    return masks[order] if len(order) else masks


def main():
    import tests.test_merge_contours as ref_tests  # the reference's fixtures (polygons)
    out = {}
    merge_seg = reference_segmenter(1024, True, 0, ['0'], ['1'])

    T = ref_tests.TestMergeContours
    cases = {'two': T.INPUT_CONTOURS_TWO_SUB_IMAGES, 'three': T.INPUT_CONTOURS_THREE_SUB_IMAGES,
             'one_empty': T.INPUT_CONTOURS_ONE_SUB_IMAGE_EMPTY, 'all_empty': T.INPUT_BOXES_ALL_SUB_IMAGES_EMPTY,
             'no_overlap2': T.INPUT_CONTOURS_NO_OVERLAP[:2], 'no_overlap3': T.INPUT_CONTOURS_NO_OVERLAP[:3]}
    size = 1024
    names = []
    for name, polys in cases.items():
        per_sub = {str(i): {'printed_text': [[contour_from_polygon(p) for p in sub]]} for i, sub in enumerate(polys)}
        names.append(name)
        out[f'merge/{name}/n_sub'] = numpy.int64(len(polys))
        for i, sub in enumerate(polys):
            out[f'merge/{name}/in{i}'] = filled(per_sub[str(i)]['printed_text'][0], size)
        for keep in (True, False):
            got = merge_seg.merge_contours_of_same_class_from_different_images(per_sub, 1, keep, ('printed_text',))['printed_text'][0]
            mine = co.merge_across_sub_images(per_sub, 1, keep, ('printed_text',))['printed_text'][0]
            assert (got is None) == (mine is None), (name, keep)
            tag = f'merge/{name}/keep{int(keep)}'
            if got is None:
                out[tag + '/none'] = numpy.int64(1)
                continue
            a, b = canonical(filled(list(got), size)), canonical(filled(list(mine), size))
            assert a.shape == b.shape and numpy.array_equal(a, b), (name, keep)
            out[tag] = a
        if name in ('two', 'three'):
            flat = {'printed_text': [[c for sub in per_sub.values() for c in sub['printed_text'][0]]]}
            got = merge_seg.merge_contours_of_same_class_from_same_image(flat)['printed_text'][0]
            mine = co.merge_within_image(flat)['printed_text'][0]
            a, b = canonical(filled(list(got), size)), canonical(filled(list(mine), size))
            assert numpy.array_equal(a, b)
            out[f'merge/{name}/same_image'] = a
    out['merge/names'] = numpy.array(names)

    # full create_segmentation_image cases
    full = [('a', 11, 6, 256, True, 10, ()), ('b', 12, 6, 256, False, 10, ()), ('c', 13, 4, 256, True, 40, ()), ('d', 14, 4, 128, True, 0, ()),
            ('e', 15, 3, 256, True, 10, (1,))]
    ids = []
    for tag, seed, batch, size, keep, min_area, blobs in full:
        pred = synthetic_document_masks(seed, batch, size, blobs)
        seg = reference_segmenter(size, keep, min_area, ['8', '9'], ['12', '13'])
        as_torch = {k: {n: torch.from_numpy(m) for n, m in v.items()} for k, v in pred.items()}
        seg.prepare_image_segmentation = types.MethodType(lambda self, activations, class_label_map: as_torch, seg)
        images, drop = seg.create_segmentation_image({0: torch.zeros(batch, 1)})
        mine_images, mine_drop = co.create_segmentation_image(pred, batch, size, seg.class_to_color_map, ['8', '9'], ['12', '13'], keep, min_area)
        assert numpy.array_equal(images, mine_images), tag
        assert sorted(drop) == sorted(mine_drop), tag
        ids.append(tag)
        out[f'full/{tag}/cfg'] = numpy.array([seed, batch, size, int(keep), min_area], dtype=numpy.int64)
        out[f'full/{tag}/blobs'] = numpy.array(list(blobs), dtype=numpy.int64)
        for k, v in pred.items():
            for n, m in v.items():
                out[f'full/{tag}/mask/{k}/{n}'] = numpy.packbits(m, axis=-1)
        out[f'full/{tag}/images'] = images
        out[f'full/{tag}/drop'] = numpy.array(sorted(drop), dtype=numpy.int64)
        print(tag, 'images', images.shape, 'coloured px', int((images.sum(-1) > 0).sum()), 'drop', sorted(drop))
    out['full/ids'] = numpy.array(ids)
    path = os.path.join(HERE, 'golden_contours_v1.npz')
    numpy.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB,', len(out), 'arrays')


if __name__ == '__main__':
    main()
