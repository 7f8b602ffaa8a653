#!/usr/bin/env python
"""Time sis_contour_stage (csrc/contours.cu) alone: B x 256^2 masks, CUDA events around the call, per input kind.
Usage: python scripts/contour_stage_bench.py [--batch 32] [--size 256] [--iters 10]   (run under ncu for the launch list)"""
import argparse
import json
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesis_in_style_b200 import contours as pc, contours_device as pd  # noqa: E402

COLORS = {'background': (0, 0, 0), 'printed_text': (0, 0, 255), 'handwritten_text': (255, 0, 0)}


def smooth_argmax_masks(batch, size, sigmas, device, seed):
    """argmax of blurred noise fields (three classes), per key: coarse blobs for the class keys, fine ones for the others."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for key, sigma in zip(('8', '9', '12', '13'), sigmas):
        x = torch.randn(batch, 3, size, size, device=device, generator=g)
        k = max(1, int(3 * sigma)) * 2 + 1
        t = torch.arange(k, device=device) - k // 2
        w = torch.exp(-t.float() ** 2 / (2 * sigma * sigma))
        w = (w / w.sum()).view(1, 1, 1, k)
        x = torch.nn.functional.conv2d(x.view(-1, 1, size, size), w, padding=(0, k // 2))
        x = torch.nn.functional.conv2d(x, w.transpose(2, 3), padding=(k // 2, 0)).view(batch, 3, size, size)
        x[:, 0] += 0.35 * x.std()            # more background than text
        ids = x.argmax(1)
        out[key] = (list(COLORS), torch.stack([(ids == j).to(torch.uint8) for j in range(3)]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--iters', type=int, default=10)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = pc.ContourConfig(args.size, COLORS, ['8', '9'], ['12', '13'], False, 50)
    for name, sigmas in (('blobs', (6, 6, 2.5, 2.5)), ('specks', (3, 3, 1, 1))):
        stacked = smooth_argmax_masks(args.batch, args.size, sigmas, dev, 7)
        stage = pd.DeviceContourStage(cfg)
        stage.run(stacked)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            out, flags = stage.run(stacked)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        f = flags.cpu().numpy()
        print(json.dumps({'input': name, 'batch': args.batch, 'size': args.size, 'ms_per_batch': round(ms, 3),
                          'images_per_s': round(args.batch / ms * 1e3), 'shapes': stage.last_info[0], 'fixpoint_rounds': stage.last_info[1],
                          'flags': {'keep': int((f == 0).sum()), 'drop': int((f == 1).sum()), 'host': int((f == 2).sum())}}), flush=True)


if __name__ == '__main__':
    main()
