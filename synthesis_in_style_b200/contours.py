"""Host-side contour post-processing of the labelled masks: class masks -> colour label image + images to drop
(the step right after the GPU hot path).

Mirrors, name for name, the contour methods of the reference's segmenters
  scf/segmentation/base_cluster_based_dataset_segmenter.py:148-450
  scf/segmentation/black_white_handwritten_printed_text_segmenter.py:42-99
  scf/segmentation/base_dataset_segmenter.py:52-57, scf/utils/segmentation_utils.py:60-85
and gives the same results (tests/test_contours.py: the reference's own merge fixtures and golden label images produced
by the reference's classes), but not with the reference's algorithms:

  * every contour is rasterised ONCE into a mask cropped to its bounding box (`Raster`); overlap tests are a strict
    bounding-box test (the reference's BBox.is_overlapping_with) plus an AND over the intersection window, instead of
    two fresh full-size canvases per test;
  * merge_contours replays the reference's merge ORDER (first overlapping pair in dict order, result appended, restart)
    without restarting: the next pair the reference merges is always the lexicographically smallest overlapping pair
    of live ids, so a priority queue of bounding-box-overlapping pairs gives the same sequence.  Every pair is tested at
    most once.  The reference is O(n^3) canvas drawings, this is one vectorised n x n box test + O(n) rasters;
  * rendering and classification work on bounding-box windows.

Everything here is per image, so `segment_masks_parallel` can fan a batch out over a process / thread pool
(OpenCV releases the GIL) while the GPU produces the next batch.  OpenCV is used for the same four primitives the
reference uses (dilate, findContours, drawContours, contourArea / boundingRect); it is a dependency of the reference
itself.
"""
from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import cv2
import numpy

Contour = numpy.ndarray
ClassContours = Dict[str, List[Optional[List[Contour]]]]
_CROSS3 = cv2.getStructuringElement(cv2.MORPH_CROSS, (3, 3)).astype(numpy.uint8)


# --------------------------------------------------------------------------- rasters

class Raster:
    """A contour with its filled drawing cropped to the bounding box [x0..x1] x [y0..y1] (inclusive pixel coords)."""
    __slots__ = ('contour', 'x0', 'y0', 'x1', 'y1', 'mask')

    def __init__(self, contour: Contour):
        self.contour = contour
        x, y, w, h = cv2.boundingRect(contour)          # = min / max of the points (segmentation_utils.py BBox of a contour)
        self.x0, self.y0, self.x1, self.y1 = x, y, x + w - 1, y + h - 1
        self.mask = numpy.zeros((self.y1 - self.y0 + 1, self.x1 - self.x0 + 1), dtype=numpy.uint8)
        cv2.drawContours(self.mask, [contour - (self.x0, self.y0)], 0, 1, cv2.FILLED)


class RasterCache:
    """Rasters of the contours seen during one create_segmentation_image call (keyed by object identity; the cache
    keeps the arrays alive, so ids are not reused)."""

    def __init__(self):
        self._by_id = {}

    def get(self, contour: Contour) -> Raster:
        r = self._by_id.get(id(contour))
        if r is None or r.contour is not contour:
            r = Raster(contour)
            self._by_id[id(contour)] = r
        return r

    def put(self, raster: Raster):
        self._by_id[id(raster.contour)] = raster


def raster_overlap(a: Raster, b: Raster) -> int:
    """contour_overlap, base_cluster_based…:156-184: strict bounding-box test, then the number of common pixels."""
    if not (a.x0 < b.x1 and a.x1 > b.x0 and a.y0 < b.y1 and a.y1 > b.y0):
        return 0
    x0, x1 = max(a.x0, b.x0), min(a.x1, b.x1)
    y0, y1 = max(a.y0, b.y0), min(a.y1, b.y1)
    wa = a.mask[y0 - a.y0:y1 - a.y0 + 1, x0 - a.x0:x1 - a.x0 + 1]
    wb = b.mask[y0 - b.y0:y1 - b.y0 + 1, x0 - b.x0:x1 - b.x0 + 1]
    return int(cv2.countNonZero(cv2.bitwise_and(wa, wb)))


def raster_union(a: Raster, b: Raster) -> Raster:
    """merge_two_contours_if_overlapping, :186-194: first external contour (CHAIN_APPROX_NONE) of the union of the two
    filled drawings, on a window one pixel larger than the joint bounding box."""
    x0, y0 = min(a.x0, b.x0), min(a.y0, b.y0)
    x1, y1 = max(a.x1, b.x1), max(a.y1, b.y1)
    canvas = numpy.zeros((y1 - y0 + 3, x1 - x0 + 3), dtype=numpy.uint8)
    for r in (a, b):
        win = canvas[r.y0 - y0 + 1:r.y1 - y0 + 2, r.x0 - x0 + 1:r.x1 - x0 + 2]
        numpy.bitwise_or(win, r.mask, out=win)
    found, _ = cv2.findContours(canvas, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    return Raster(found[0] + (x0 - 1, y0 - 1))


# --------------------------------------------------------------------------- reference primitives, same names

def dilate_image(image: numpy.ndarray, kernel: numpy.ndarray = None, kernel_size: int = 3) -> numpy.ndarray:
    """base_dataset_segmenter.py:52-57."""
    if kernel is None:
        kernel = _CROSS3 if kernel_size == 3 else cv2.getStructuringElement(cv2.MORPH_CROSS, (kernel_size, kernel_size)).astype(numpy.uint8)
    return cv2.morphologyEx(image, cv2.MORPH_DILATE, kernel)


def cluster_image_to_contours(cluster_arrays: numpy.ndarray) -> List[Sequence[Contour]]:
    """:148-154."""
    out = []
    for image in cluster_arrays:
        found, _ = cv2.findContours(dilate_image(image), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        out.append(found)
    return out


def contour_overlap(contour1: Contour, contour2: Contour, cache: Optional[RasterCache] = None) -> int:
    cache = cache or RasterCache()
    return raster_overlap(cache.get(contour1), cache.get(contour2))


def merge_two_contours_if_overlapping(contour1: Contour, contour2: Contour, cache: Optional[RasterCache] = None):
    cache = cache or RasterCache()
    a, b = cache.get(contour1), cache.get(contour2)
    if raster_overlap(a, b) <= 0:
        return None
    merged = raster_union(a, b)
    cache.put(merged)
    return (merged.contour,)


def merge_contours(contours: Sequence[Contour], only_keep_overlapping: bool = False,
                   cache: Optional[RasterCache] = None) -> List[Contour]:
    """merge_contours / _try_merge_contours, :196-226, same result AND same order, without the restarts.

    The reference repeats {scan the pairs of the dict in order, merge the first overlapping one, append the union,
    start over}.  Dict order is ascending id (unions get the next id), a pair that did not overlap never will (its two
    shapes do not change), so the pair it merges next is always the lexicographically smallest overlapping pair of live
    ids.  That is a priority queue: seed it with the bounding-box-overlapping pairs (one vectorised n x n test), pop in
    order, test the rasters, and after a merge push the new contour's bounding-box partners."""
    import heapq
    cache = cache or RasterCache()
    n = len(contours)
    if n == 0:
        return []
    # bounding boxes first (cheap); a contour is rasterised only when a pair it belongs to is actually tested, so the
    # isolated specks that dominate fine-grained masks never are
    box = numpy.empty((2 * n, 4), dtype=numpy.int64)
    for i, c in enumerate(contours):
        x, y, w, h = cv2.boundingRect(c)
        box[i] = (x, y, x + w - 1, y + h - 1)
    rasters: List[Optional[Raster]] = [None] * n
    shapes: List[Contour] = list(contours)

    def raster(i: int) -> Raster:
        if rasters[i] is None:
            rasters[i] = cache.get(shapes[i])
        return rasters[i]

    alive = [True] * n
    members = [1] * n
    b = box[:n]
    touching = (b[:, None, 0] < b[None, :, 2]) & (b[:, None, 2] > b[None, :, 0]) \
        & (b[:, None, 1] < b[None, :, 3]) & (b[:, None, 3] > b[None, :, 1])
    heap = [tuple(p) for p in numpy.argwhere(numpy.triu(touching, 1)).tolist()]      # sorted, hence already a heap
    while heap:
        i, j = heapq.heappop(heap)
        if not (alive[i] and alive[j]):
            continue
        if raster_overlap(raster(i), raster(j)) <= 0:
            continue
        merged = raster_union(raster(i), raster(j))
        cache.put(merged)
        k = len(shapes)
        shapes.append(merged.contour)
        rasters.append(merged)
        box[k] = (merged.x0, merged.y0, merged.x1, merged.y1)
        alive[i] = alive[j] = False
        alive.append(True)
        members.append(members[i] + members[j])
        bk = box[:k]
        partners = numpy.flatnonzero((bk[:, 0] < merged.x1) & (bk[:, 2] > merged.x0) & (bk[:, 1] < merged.y1) & (bk[:, 3] > merged.y0))
        for q in partners.tolist():
            if alive[q]:
                heapq.heappush(heap, (q, k))
    return [shapes[i] for i in range(len(shapes)) if alive[i] and (members[i] > 1 or not only_keep_overlapping)]


# --------------------------------------------------------------------------- batch-level methods (reference names)

def merge_contours_of_same_class_from_different_images(class_contours_for_sub_images, batch_size: int,
                                                       only_keep_overlapping: bool = False,
                                                       class_names_to_merge: Tuple[str, ...] = (),
                                                       drop_if_size_of_contours_zero: bool = False,
                                                       cache: Optional[RasterCache] = None) -> ClassContours:
    """:228-302."""
    cache = cache or RasterCache()
    if len(class_names_to_merge) == 0:
        class_names_to_merge = {name for sub in class_contours_for_sub_images.values() for name in sub.keys()}
    by_class = defaultdict(list)
    for sub in class_contours_for_sub_images.values():
        for name, batches in sub.items():
            by_class[name].append(batches)
    result = defaultdict(list)
    for name, per_sub in by_class.items():
        mergeable = name in class_names_to_merge
        for b in range(batch_size):
            here = [batches[b] for batches in per_sub]
            empty = [len(c) == 0 for c in here]
            if all(empty) or (drop_if_size_of_contours_zero and mergeable and any(empty)):
                result[name].append(None)
            elif any(empty):
                result[name].append(here[empty.index(False)])
            else:
                flat = [c for sub_contours in here for c in sub_contours]
                if not mergeable or len(here) == 1:
                    result[name].append(flat)
                else:
                    merged = merge_contours(flat, only_keep_overlapping, cache)
                    result[name].append(merged if len(merged) else None)
    return result


def merge_contours_of_same_class_from_same_image(class_contours: ClassContours, cache: Optional[RasterCache] = None) -> ClassContours:
    """:304-316."""
    cache = cache or RasterCache()
    return {name: [None if c is None else merge_contours(c, False, cache) for c in batches]
            for name, batches in class_contours.items()}


def extract_contours(predicted_clusters, image_ids_to_extract: Sequence[str]):
    """:318-332; masks may be torch tensors (any device) or numpy arrays [B,S,S]."""
    out = {}
    for key in image_ids_to_extract:
        per_class = {}
        for name, mask in predicted_clusters[key].items():
            if name == 'background':
                continue
            per_class[name] = cluster_image_to_contours(_as_uint8(mask))
        out[key] = per_class
    return out


def drop_too_small_contours(class_contours: ClassContours, min_class_contour_area: float) -> ClassContours:
    """:393-405."""
    out = {}
    for name, batches in class_contours.items():
        kept = []
        for contours in batches:
            if contours is not None:
                contours = [c for c in contours if cv2.contourArea(c) >= min_class_contour_area]
                if len(contours) == 0:
                    contours = None
            kept.append(contours)
        out[name] = kept
    return out


def classify_fine_grained_contours(text_regions_per_class: ClassContours, fine_grained_contours_per_class: ClassContours,
                                   class_id_map: Dict[str, int], fine_grained_class_name: str = 'printed_text',
                                   cache: Optional[RasterCache] = None) -> ClassContours:
    """:351-391: every fine-grained contour goes to the class whose text regions it overlaps most."""
    assert len(text_regions_per_class) == len(fine_grained_contours_per_class), \
        'Num classes of text regions and fine grained contours must be equal! '
    cache = cache or RasterCache()
    fine_batches = fine_grained_contours_per_class[fine_grained_class_name]
    names = sorted(text_regions_per_class.keys(), key=lambda n: class_id_map[n])
    batch_size = len(fine_batches)
    classified = {name: [] for name in names}
    for b in range(batch_size):
        fine = fine_batches[b]
        picked = {name: [] for name in names}
        if fine is not None and len(fine) > 0:
            live = [name for name in names if text_regions_per_class[name][b] is not None]
            if live:
                fine_r = [cache.get(c) for c in fine]
                region_r = {name: [cache.get(c) for c in text_regions_per_class[name][b]] for name in live}
                for cid, fr in enumerate(fine_r):
                    best, best_score = None, 0
                    for name in live:                         # colour-map order; strict > keeps the first maximum
                        score = 0
                        for rr in region_r[name]:
                            score += raster_overlap(fr, rr)
                        if score > best_score:
                            best, best_score = name, score
                    if best is not None:
                        picked[best].append(fine[cid])
        for name in names:
            classified[name].append(picked[name] if picked[name] else None)
    return classified


def bounding_rect_from_contours(contours: Sequence[Contour]) -> numpy.ndarray:
    """segmentation_utils.py:60-64, including its shape: ONE row of 4n numbers (x, y, w, h, x, y, w, h, ...)."""
    rects = numpy.concatenate([cv2.boundingRect(c) for c in contours])
    if rects.ndim == 1:
        rects = rects.reshape((1, len(rects)))
    return rects


def determine_images_to_drop(fine_grained_contours_per_image: ClassContours, image_size: int) -> List[int]:
    """black_white…:61-75.  With the one-row rect array above, columns 2 and 3 are the width and height of the FIRST
    contour of each class: that is the rule the reference applies, so it is the rule applied here."""
    drop = set()
    limit = int(image_size * 0.95)
    for batches in fine_grained_contours_per_image.values():
        for image_id, contours in enumerate(batches):
            if contours is None:
                continue
            rects = bounding_rect_from_contours(contours)
            if (rects[:, 3] > limit).any() and (rects[:, 2] > limit).any():
                drop.add(image_id)
    return list(drop)


def render_segmentation_image(fine_grained_prediction, classified_contours: ClassContours, batch_size: int, image_size: int,
                              class_to_color_map: Dict[str, Tuple[int, int, int]], cluster_class_name: str = 'printed_text',
                              cache: Optional[RasterCache] = None) -> numpy.ndarray:
    """:407-447 on bounding-box windows: inside each classified contour, the pixels of `cluster_class_name`'s mask of
    the last fine-grained key take the class colour (classes in mask-dict order, later ones overwrite)."""
    cache = cache or RasterCache()
    ink = _as_uint8(fine_grained_prediction[cluster_class_name])
    out = numpy.empty((batch_size, image_size, image_size, 3), dtype=numpy.uint8)
    out[:] = numpy.asarray(class_to_color_map['background'], dtype=numpy.uint8)
    for b in range(batch_size):
        for name in fine_grained_prediction.keys():
            if name == 'background':
                continue
            contours = classified_contours[name][b]
            if contours is None:
                continue
            color = numpy.asarray(class_to_color_map[name], dtype=numpy.uint8)
            for contour in contours:
                r = cache.get(contour)
                # cv2 clips drawings to the canvas; contours come from S x S masks, so the window is inside the image
                x0, y0 = max(r.x0, 0), max(r.y0, 0)
                x1, y1 = min(r.x1, image_size - 1), min(r.y1, image_size - 1)
                if x1 < x0 or y1 < y0:
                    continue
                inside = r.mask[y0 - r.y0:y1 - r.y0 + 1, x0 - r.x0:x1 - r.x0 + 1]
                sel = (inside != 0) & (ink[b, y0:y1 + 1, x0:x1 + 1] != 0)
                out[b, y0:y1 + 1, x0:x1 + 1][sel] = color
    return out


def _as_uint8(mask) -> numpy.ndarray:
    if hasattr(mask, 'detach'):
        mask = mask.detach().cpu().numpy()
    mask = numpy.asarray(mask)
    return mask if mask.dtype == numpy.uint8 else mask.astype(numpy.uint8)


# --------------------------------------------------------------------------- driver

class ContourConfig:
    """The creation-config keys the contour stage reads (configs/dataset_creation/*.json)."""

    def __init__(self, image_size: int, class_to_color_map: Dict[str, Tuple[int, int, int]],
                 keys_for_class_determination: Sequence[str], keys_for_finegrained_segmentation: Sequence[str],
                 only_keep_overlapping: bool = True, min_class_contour_area: float = 0):
        self.image_size = image_size
        self.class_to_color_map = dict(class_to_color_map)
        self.class_id_map = {name: i for i, name in enumerate(self.class_to_color_map)}
        self.keys_for_class_determination = list(keys_for_class_determination)
        self.keys_for_finegrained_segmentation = list(keys_for_finegrained_segmentation)
        self.only_keep_overlapping = only_keep_overlapping
        self.min_class_contour_area = min_class_contour_area


def segment_masks(predicted_clusters, batch_size: int, cfg: ContourConfig):
    """create_segmentation_image, black_white…:77-99, from the merged PredictedClusters on: (uint8 [B,S,S,3], drop ids)."""
    cache = RasterCache()
    host = {key: {name: _as_uint8(mask) for name, mask in per_class.items()} for key, per_class in predicted_clusters.items()
            if key in cfg.keys_for_class_determination or key in cfg.keys_for_finegrained_segmentation}
    # extract_text_regions, black_white…:42-59
    regions = merge_contours_of_same_class_from_different_images(
        extract_contours(host, cfg.keys_for_class_determination), batch_size,
        only_keep_overlapping=cfg.only_keep_overlapping, drop_if_size_of_contours_zero=True, cache=cache)
    regions = drop_too_small_contours(regions, cfg.min_class_contour_area)
    # merge_finegrained_segmentation, base_cluster_based…:334-349
    fine = merge_contours_of_same_class_from_different_images(
        extract_contours(host, cfg.keys_for_finegrained_segmentation), batch_size,
        only_keep_overlapping=True, drop_if_size_of_contours_zero=True, cache=cache)
    classified = classify_fine_grained_contours(regions, fine, cfg.class_id_map, 'printed_text', cache=cache)
    classified = drop_too_small_contours(classified, cfg.min_class_contour_area)
    drop = determine_images_to_drop(classified, cfg.image_size)
    images = render_segmentation_image(host[cfg.keys_for_finegrained_segmentation[-1]], classified, batch_size,
                                       cfg.image_size, cfg.class_to_color_map, 'printed_text', cache=cache)
    return images, drop


def _segment_one(args):
    masks, cfg = args
    cv2.setNumThreads(1)            # pool workers: one image per task, OpenCV's own thread pool only oversubscribes
    image, drop = segment_masks(masks, 1, cfg)
    return image[0], bool(drop)


def segment_masks_parallel(predicted_clusters, batch_size: int, cfg: ContourConfig, pool=None):
    """Same result as segment_masks, one task per image on `pool` (a concurrent.futures executor; images are
    independent: every step of the reference loops over the batch).  Without a pool it is segment_masks."""
    if pool is None or batch_size <= 1:
        return segment_masks(predicted_clusters, batch_size, cfg)
    keys = set(cfg.keys_for_class_determination) | set(cfg.keys_for_finegrained_segmentation)
    host = {key: {name: _as_uint8(mask) for name, mask in per_class.items()}
            for key, per_class in predicted_clusters.items() if key in keys}
    tasks = [({key: {name: m[b:b + 1] for name, m in per_class.items()} for key, per_class in host.items()}, cfg)
             for b in range(batch_size)]
    results = list(pool.map(_segment_one, tasks))
    images = numpy.stack([r[0] for r in results], axis=0)
    return images, [b for b, r in enumerate(results) if r[1]]
