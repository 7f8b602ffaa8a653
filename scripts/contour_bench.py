#!/usr/bin/env python
"""CPU timing of the contour stage on synthetic document-like masks (the generator of
tests/golden/make_golden_contours.py): product (synthesis_in_style_b200/contours.py) vs the restatement of the
reference's algorithm (oracle/contour_oracle.py), same inputs, results asserted equal.  No GPU needed.
Usage: scripts/contour_bench.py [--images 16] [--oracle-images 4] [--workers 8]"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import numpy  # noqa: E402

from make_golden_contours import synthetic_document_masks  # noqa: E402
from oracle import contour_oracle as co  # noqa: E402
from synthesis_in_style_b200 import contours as pc  # noqa: E402

COLORS = {'background': (0, 0, 0), 'printed_text': (0, 0, 255), 'handwritten_text': (255, 0, 0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=16)
    ap.add_argument('--oracle-images', type=int, default=4)
    ap.add_argument('--workers', type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument('--size', type=int, default=256)
    a = ap.parse_args()
    pred = synthetic_document_masks(21, a.images, a.size)
    cfg = pc.ContourConfig(a.size, COLORS, ['8', '9'], ['12', '13'], True, 10)
    pc.segment_masks({k: {n: m[:1] for n, m in v.items()} for k, v in pred.items()}, 1, cfg)     # warm-up
    t0 = time.perf_counter()
    images, drop = pc.segment_masks(pred, a.images, cfg)
    t_prod = time.perf_counter() - t0
    with ProcessPoolExecutor(a.workers) as pool:
        pc.segment_masks_parallel(pred, a.images, cfg, pool)                                      # spawn + warm-up
        t0 = time.perf_counter()
        images_p, drop_p = pc.segment_masks_parallel(pred, a.images, cfg, pool)
        t_par = time.perf_counter() - t0
    assert numpy.array_equal(images, images_p) and sorted(drop) == sorted(drop_p)
    n = a.oracle_images
    sub = {k: {nm: m[:n] for nm, m in v.items()} for k, v in pred.items()}
    t0 = time.perf_counter()
    o_images, o_drop = co.create_segmentation_image(sub, n, a.size, COLORS, ['8', '9'], ['12', '13'], True, 10)
    t_or = time.perf_counter() - t0
    assert numpy.array_equal(o_images, images[:n]) and sorted(o_drop) == sorted(d for d in drop if d < n)
    print(json.dumps({'stage': 'contours', 'image_size': a.size, 'images': a.images,
                      'product_ms_per_image_1_core': round(t_prod / a.images * 1e3, 2),
                      'product_images_per_s_pool': round(a.images / t_par, 1), 'pool_workers': a.workers,
                      'reference_algorithm_ms_per_image_1_core': round(t_or / n * 1e3, 1),
                      'speedup_1_core': round((t_or / n) / (t_prod / a.images), 1), 'results_equal': True}))


if __name__ == '__main__':
    main()
