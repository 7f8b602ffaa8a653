"""CPU restatement of the reference's contour post-processing: class masks ->
colour label image + list of images to drop.  TEST INFRASTRUCTURE ONLY (tests/, bench cpu_baseline leg).

It follows, step by step and with the same OpenCV calls (cv2 is a dependency of the reference itself):
  scf/segmentation/base_dataset_segmenter.py:52-57                 dilate_image
  scf/segmentation/base_cluster_based_dataset_segmenter.py:148-450 contours, overlap, merging, classification, render
  scf/segmentation/black_white_handwritten_printed_text_segmenter.py:42-99  text regions, drop rule, driver
  scf/utils/segmentation_utils.py:22-85                             BBox, bounding rects, canvases
Pinned by tests/golden/golden_contours_v1.npz, which tests/golden/make_golden_contours.py generates by running the
reference's own classes in this container (and by the reference's tests/test_merge_contours.py fixtures).
The algorithms are deliberately the reference's slow ones (restart-after-every-merge pair search, a fresh canvas per
overlap test); the product implementation in synthesis_in_style_b200/contours.py must give the same results faster.
"""
from collections import defaultdict
from itertools import combinations
from typing import Dict, List, Optional, Sequence, Tuple

import cv2
import numpy

Contour = numpy.ndarray                                  # [n_points, 1, 2] int32, (x, y)
ClassContours = Dict[str, List[Optional[List[Contour]]]]  # class -> per image: contours or None


# --------------------------------------------------------------------------- primitives

def dilate_image(image: numpy.ndarray, kernel_size: int = 3) -> numpy.ndarray:
    """base_dataset_segmenter.py:52-57: 3x3 cross dilation."""
    kernel = cv2.getStructuringElement(cv2.MORPH_CROSS, (kernel_size, kernel_size)).astype(numpy.uint8)
    return cv2.morphologyEx(image, cv2.MORPH_DILATE, kernel)


def masks_to_contours(masks: numpy.ndarray) -> List[Sequence[Contour]]:
    """cluster_image_to_contours, base_cluster_based…:148-154: per image, dilate then external contours."""
    out = []
    for image in masks:
        found, _ = cv2.findContours(dilate_image(image), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        out.append(found)
    return out


def bbox_of(contour: Contour) -> Tuple[int, int, int, int]:
    (x0, y0), (x1, y1) = contour.min(axis=0)[0], contour.max(axis=0)[0]
    return int(x0), int(y0), int(x1), int(y1)


def bboxes_overlap(a, b) -> bool:
    """BBox.is_overlapping_with, segmentation_utils.py:50-54 (strict: touching boxes do not overlap)."""
    return a[0] < b[2] and a[2] > b[0] and a[1] < b[3] and a[3] > b[1]


def canvases(contours: Sequence[Contour], minimal: bool = False) -> List[numpy.ndarray]:
    """draw_contours_on_same_sized_canvases, segmentation_utils.py:71-85: one filled drawing per contour."""
    pts = numpy.concatenate(contours)
    x_max, y_max = pts.max(axis=0)[0]
    x_min, y_min = (pts.min(axis=0)[0] if minimal else (0, 0))
    blank = numpy.zeros((y_max - y_min + 1, x_max - x_min + 1))
    return [cv2.drawContours(blank.copy(), [c - (x_min, y_min)], 0, 1, cv2.FILLED) for c in contours]


def contour_overlap(a: Contour, b: Contour) -> int:
    """contour_overlap, base_cluster_based…:156-184: number of pixels both filled drawings cover."""
    if not bboxes_overlap(bbox_of(a), bbox_of(b)):
        return 0
    first, second = canvases([a, b], minimal=True)
    return int(numpy.logical_and(first, second).sum())


def merge_pair(a: Contour, b: Contour):
    """merge_two_contours_if_overlapping, :186-194: external contours (CHAIN_APPROX_NONE) of the union, or None."""
    if contour_overlap(a, b) <= 0:
        return None
    first, second = canvases([a, b])
    union = numpy.logical_or(first, second).astype(numpy.uint8) * 255
    found, _ = cv2.findContours(union, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    return found


def merge_contours(contours: Sequence[Contour], only_keep_overlapping: bool = False) -> List[Contour]:
    """merge_contours / _try_merge_contours, :196-226: repeat {first overlapping pair in combinations order -> replace
    both by the FIRST contour of their union, appended at the end} until no pair overlaps."""
    pool = {(i,): contours[i] for i in range(len(contours))}
    merged_something = True
    while merged_something:
        merged_something = False
        for ka, kb in combinations(list(pool.keys()), 2):
            union = merge_pair(pool[ka], pool[kb])
            if union is not None:
                pool[ka + kb] = union[0]
                del pool[ka], pool[kb]
                merged_something = True
                break
    if only_keep_overlapping:
        return [c for ids, c in pool.items() if len(ids) > 1]
    return list(pool.values())


# --------------------------------------------------------------------------- per-class / per-image logic

def merge_across_sub_images(per_sub_image: Dict[str, Dict[str, List[Sequence[Contour]]]], batch_size: int,
                            only_keep_overlapping: bool = False, class_names_to_merge: Tuple[str, ...] = (),
                            drop_if_size_of_contours_zero: bool = False) -> ClassContours:
    """merge_contours_of_same_class_from_different_images, :228-302."""
    if len(class_names_to_merge) == 0:
        class_names_to_merge = {name for sub in per_sub_image.values() for name in sub.keys()}
    by_class = defaultdict(list)
    for sub in per_sub_image.values():
        for name, batches in sub.items():
            by_class[name].append(batches)
    result = defaultdict(list)
    for name, per_sub in by_class.items():
        for b in range(batch_size):
            here = [batches[b] for batches in per_sub]
            empty = [len(c) == 0 for c in here]
            if all(empty):
                result[name].append(None)
                continue
            if drop_if_size_of_contours_zero and name in class_names_to_merge and any(empty):
                result[name].append(None)
                continue
            if any(empty):
                for i, is_empty in enumerate(empty):
                    if not is_empty:
                        result[name].append(per_sub[i][b])
                        break
                continue
            flat = [c for sub_contours in here for c in sub_contours]
            if name not in class_names_to_merge or len(here) == 1:
                result[name].append(flat)
                continue
            merged = merge_contours(flat, only_keep_overlapping)
            result[name].append(merged if len(merged) else None)
    return result


def merge_within_image(class_contours: ClassContours) -> ClassContours:
    """merge_contours_of_same_class_from_same_image, :304-316."""
    return {name: [None if c is None else merge_contours(c) for c in batches] for name, batches in class_contours.items()}


def extract_contours(predicted: Dict[str, Dict[str, numpy.ndarray]], keys: Sequence[str]):
    """extract_contours, :318-332 (masks as uint8/bool arrays [B,S,S]; 'background' skipped)."""
    out = {}
    for key in keys:
        out[key] = {name: masks_to_contours(numpy.asarray(mask).astype(numpy.uint8))
                    for name, mask in predicted[key].items() if name != 'background'}
    return out


def drop_too_small(class_contours: ClassContours, min_area: float) -> ClassContours:
    """drop_too_small_contours, :393-405 (cv2.contourArea = polygon area of the point chain, not a pixel count)."""
    out = {}
    for name, batches in class_contours.items():
        kept = []
        for contours in batches:
            if contours is not None:
                contours = [c for c in contours if cv2.contourArea(c) >= min_area]
                if len(contours) == 0:
                    contours = None
            kept.append(contours)
        out[name] = kept
    return out


def classify_fine_grained(text_regions: ClassContours, fine_grained: ClassContours, class_id_map: Dict[str, int],
                          fine_grained_class_name: str = 'printed_text') -> ClassContours:
    """classify_fine_grained_contours, :351-391: each fine-grained contour goes to the class whose text regions it
    overlaps most (first class in colour-map order on ties, dropped when every overlap is 0)."""
    assert len(text_regions) == len(fine_grained)
    classified = defaultdict(list)
    fine_batches = fine_grained[fine_grained_class_name]
    text_regions = dict(sorted(text_regions.items(), key=lambda kv: class_id_map[kv[0]]))
    batch_size = len(fine_batches)
    ranking = {i: defaultdict(dict) for i in range(batch_size)}
    for name, region_batches in text_regions.items():
        for b, (regions, fine) in enumerate(zip(region_batches, fine_batches)):
            if regions is None or fine is None or len(fine) == 0:
                classified[name].append(None)
                continue
            for cid, contour in enumerate(fine):
                scores = ranking[b][cid]
                if name not in scores:
                    scores[name] = 0
                for region in regions:
                    scores[name] += contour_overlap(contour, region)
    for name in text_regions.keys():
        classified[name] = [[] for _ in range(batch_size)]
    for b in range(batch_size):
        for cid, scores in ranking[b].items():
            best = max(scores, key=scores.get)
            if scores[best] > 0:
                classified[best][b].append(fine_batches[b][cid])
        for name in text_regions.keys():
            if len(classified[name][b]) == 0:
                classified[name][b] = None
    return classified


def images_to_drop(classified: ClassContours, image_size: int) -> List[int]:
    """determine_images_to_drop, black_white…:61-75, with bounding_rect_from_contours, segmentation_utils.py:60-64.
    Reference quirk kept: concatenating the (x, y, w, h) tuples gives ONE row of 4n numbers (the `ndim == 1` reshape
    always fires), so columns 2 and 3 are the width and height of the FIRST contour of the class only: an image is
    dropped when that contour's bounding rect is both wider and taller than 95 % of the image."""
    drop = set()
    limit = int(image_size * 0.95)
    for batches in classified.values():
        for image_id, contours in enumerate(batches):
            if contours is None:
                continue
            rects = numpy.concatenate([cv2.boundingRect(c) for c in contours])
            rects = rects.reshape((1, len(rects)))
            if (rects[:, 3] > limit).any() and (rects[:, 2] > limit).any():
                drop.add(image_id)
    return list(drop)


def render(fine_prediction: Dict[str, numpy.ndarray], classified: ClassContours, batch_size: int, image_size: int,
           class_to_color: Dict[str, Tuple[int, int, int]], cluster_class_name: str = 'printed_text') -> numpy.ndarray:
    """render_segmentation_image, :407-447: inside every classified contour the pixels of the LAST fine-grained key's
    `printed_text` mask take the contour's class colour."""
    fine_prediction = {name: numpy.asarray(m) for name, m in fine_prediction.items()}
    images = []
    for b in range(batch_size):
        canvas = numpy.zeros((image_size, image_size, 3), dtype=numpy.uint8)
        canvas[:, :] = class_to_color['background']
        for name in fine_prediction.keys():
            if name == 'background':
                continue
            contours = classified[name][b]
            if contours is None:
                continue
            for contour in contours:
                inside = cv2.drawContours(numpy.zeros((image_size, image_size)), [contour], 0, 1, cv2.FILLED).astype(bool)
                canvas[numpy.where(inside, fine_prediction[cluster_class_name][b], False)] = class_to_color[name]
        images.append(canvas)
    return numpy.stack(images, axis=0)


# --------------------------------------------------------------------------- driver

def create_segmentation_image(predicted: Dict[str, Dict[str, numpy.ndarray]], batch_size: int, image_size: int,
                              class_to_color: Dict[str, Tuple[int, int, int]], keys_for_class_determination: Sequence[str],
                              keys_for_finegrained_segmentation: Sequence[str], only_keep_overlapping: bool,
                              min_class_contour_area: float):
    """create_segmentation_image, black_white…:77-99, from the merged PredictedClusters on (already resized, bool [B,S,S])."""
    class_id_map = {name: i for i, name in enumerate(class_to_color)}
    # extract_text_regions, black_white…:42-59
    regions = merge_across_sub_images(extract_contours(predicted, keys_for_class_determination), batch_size,
                                      only_keep_overlapping=only_keep_overlapping, drop_if_size_of_contours_zero=True)
    regions = drop_too_small(regions, min_class_contour_area)
    # merge_finegrained_segmentation, base_cluster_based…:334-349
    fine = merge_across_sub_images(extract_contours(predicted, keys_for_finegrained_segmentation), batch_size,
                                   only_keep_overlapping=True, drop_if_size_of_contours_zero=True)
    classified = classify_fine_grained(regions, fine, class_id_map, 'printed_text')
    classified = drop_too_small(classified, min_class_contour_area)
    drop = images_to_drop(classified, image_size)
    images = render(predicted[keys_for_finegrained_segmentation[-1]], classified, batch_size, image_size, class_to_color)
    return images, drop


# --------------------------------------------------------------------------- synthetic inputs (tests, bench legs)

MASK_CLASSES = ('background', 'printed_text', 'handwritten_text')


def synthetic_document_masks(seed, batch, size, blob_images=()):
    """Masks that look like the labeller's output on documents: coarse blocky text regions for the two
    class-determination keys (64^2 maps upsampled x4), stroke-like fine-grained masks at full resolution."""
    rng = numpy.random.RandomState(seed)
    coarse = size // 4
    pred = {k: {n: numpy.zeros((batch, size, size), dtype=bool) for n in MASK_CLASSES} for k in ('8', '9', '12', '13')}
    for b in range(batch):
        n_regions = rng.randint(0, 6)
        for _ in range(n_regions):
            cls = 'printed_text' if rng.rand() < 0.6 else 'handwritten_text'
            w, h = rng.randint(6, coarse // 2), rng.randint(2, 10)
            x, y = rng.randint(0, coarse - w), rng.randint(0, coarse - h)
            for key in ('8', '9'):
                if rng.rand() < 0.12:
                    continue                                             # a region one key misses
                dx, dy = rng.randint(-2, 3), rng.randint(-1, 2)
                blk = numpy.zeros((coarse, coarse), dtype=bool)
                blk[max(0, y + dy):y + dy + h, max(0, x + dx):x + dx + w] = True
                holes = rng.rand(coarse, coarse) < 0.04
                blk &= ~holes
                pred[key][cls][b] |= numpy.kron(blk, numpy.ones((4, 4), dtype=bool))
            # strokes inside the region (fine-grained keys label all ink as printed_text; a few handwritten pixels)
            for key in ('12', '13'):
                for line in range(y * 4 + 2, (y + h) * 4 - 2, 7):
                    xx = x * 4 + rng.randint(0, 4)
                    while xx < (x + w) * 4 - 3:
                        ww = rng.randint(2, 9)
                        hh = rng.randint(2, 5)
                        jit = rng.randint(-1, 2) if key == '13' else 0
                        pred[key]['printed_text'][b, max(0, line + jit):line + jit + hh, xx:min(size, xx + ww)] = True
                        xx += ww + rng.randint(1, 4)
        # specks, and occasionally a page-sized blob that triggers the drop rule
        for key in ('12', '13'):
            specks = rng.rand(size, size) < 0.0015
            pred[key]['printed_text'][b] |= specks
        if rng.rand() < 0.15 or b in blob_images:
            for key in ('8', '9', '12', '13'):
                pred[key]['printed_text'][b, 2:size - 2, 2:size - 2] |= rng.rand(size - 4, size - 4) < 0.9
        for key in pred:
            any_text = pred[key]['printed_text'][b] | pred[key]['handwritten_text'][b]
            pred[key]['handwritten_text'][b] &= ~pred[key]['printed_text'][b]
            pred[key]['background'][b] = ~any_text
    return pred
