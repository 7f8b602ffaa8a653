#!/usr/bin/env python
"""Randomised cross-check of the device contour stage (sis_contour_stage) against the host polygon path
(synthesis_in_style_b200/contours.py) on many mask statistics, image sizes (also not multiples of 32), key counts and
configurations.  Prints one line per case and a summary; exits 1 on any mismatch.
Usage: python scripts/stress_contours.py [--cases 60] [--seed 0]"""
import argparse
import os
import sys
import time

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesis_in_style_b200 import contours as pc, contours_device as pd  # noqa: E402

COLORS = {'background': (0, 0, 0), 'printed_text': (0, 0, 255), 'handwritten_text': (255, 0, 0)}


def random_masks(rng, batch, size, keys):
    from scipy import ndimage
    kind = rng.integers(0, 4)
    out = {}
    for key in keys:
        if kind == 0:      # smoothed argmax fields
            s = float(rng.choice([0.7, 1.5, 3.0, 5.0]))
            f = ndimage.gaussian_filter(rng.standard_normal((batch, 3, size, size)), (0, 0, s, s))
            f[:, 0] += rng.choice([0.0, 0.3, 0.8]) * f.std()
            ids = f.argmax(1)
        elif kind == 1:    # sparse specks
            ids = (rng.random((batch, size, size)) < rng.choice([0.01, 0.05, 0.2])).astype(int) * rng.integers(1, 3, (batch, size, size))
        elif kind == 2:    # rectangles, frames and lines
            ids = numpy.zeros((batch, size, size), int)
            for b in range(batch):
                for _ in range(int(rng.integers(1, 12))):
                    x, y = rng.integers(0, size - 2, 2)
                    w, h = rng.integers(1, max(2, size // 2), 2)
                    c = int(rng.integers(1, 3))
                    if rng.random() < 0.4:      # frame
                        t = int(rng.integers(1, 4))
                        ids[b, y:y + h, x:x + t] = c; ids[b, y:y + h, x + w - t:x + w] = c
                        ids[b, y:y + t, x:x + w] = c; ids[b, y + h - t:y + h, x:x + w] = c
                    else:
                        ids[b, y:y + h, x:x + w] = c
        else:              # mostly full
            ids = (rng.random((batch, size, size)) < 0.97).astype(int) * int(rng.integers(1, 3))
        out[key] = {name: (ids == j).astype(numpy.uint8) for j, name in enumerate(COLORS)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=60)
    ap.add_argument('--seed', type=int, default=0)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    rng = numpy.random.default_rng(args.seed)
    bad = host_flagged = unsettled = images = 0
    t0 = time.time()
    for case in range(args.cases):
        size = int(rng.choice([33, 48, 64, 100, 128, 160, 256]))
        batch = int(rng.integers(1, 5))
        n_det, n_fine = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        keys = [str(k) for k in range(n_det + n_fine)]
        det, fine = keys[:n_det], keys[n_det:]
        if rng.random() < 0.3 and n_det > 1:
            fine = fine[:-1] + [det[0]]                      # a key used by both stages
        cfg = pc.ContourConfig(size, COLORS, det, fine, bool(rng.integers(0, 2)), float(rng.choice([0, 3, 10, 50])))
        pred = random_masks(rng, batch, size, sorted(set(det + fine)))
        want_images, want_drop = pc.segment_masks(pred, batch, cfg)
        stacked = {k: (list(v), torch.stack([torch.from_numpy(m) for m in v.values()]).to(dev)) for k, v in pred.items()}
        stage = pd.DeviceContourStage(cfg)
        out, flags = stage.run(stacked)
        out, flags, info = out.cpu().numpy(), flags.cpu().numpy(), stage.last_info
        ok = True
        if info[2] == 3:
            unsettled += 1
            ok = bool((flags == 2).all())
        elif info[2] != 0:
            ok = False
        else:
            for b in range(batch):
                if not numpy.array_equal(out[b], want_images[b]):
                    ok = False
                if flags[b] != 2 and (flags[b] == 1) != (b in want_drop):
                    ok = False
        got_images, got_drop = pd.segment(stacked, batch, cfg, stage)
        ok = ok and numpy.array_equal(got_images, want_images) and got_drop == sorted(want_drop)
        images += batch
        host_flagged += int((flags == 2).sum())
        bad += not ok
        print(f'case {case}: size {size} batch {batch} keys {det}/{fine} keep {cfg.only_keep_overlapping} min_area {cfg.min_class_contour_area} '
              f'shapes {info[0]} rounds {info[1]} reason {info[2]} flags {flags.tolist()} -> {"ok" if ok else "MISMATCH"}', flush=True)
    print(f'{args.cases} cases, {images} images, {bad} mismatching cases, {host_flagged} images flagged for the host, '
          f'{unsettled} unsettled batches, {time.time() - t0:.0f} s')
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
