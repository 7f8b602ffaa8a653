#!/usr/bin/env python
"""Throughput of the DatasetGAN labeller at the BASELINE shape: 256^2, 14 captures (F = 5888),
3 networks, 3 classes.  Captures come from the B200 generator; timed: sis_pixel_ensemble_label over a batch (CUDA events),
its per-category split, and -- as the reference-on-this-GPU figure -- the reference's algorithm in PyTorch on the same
device (materialised [B,S,S,F] features, three fp32 MLPs, TF32 off) on a smaller batch.
Usage: scripts/dataset_gan_bench.py [--batch 32] [--ref-batch 2]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dataset_gan_oracle as dg  # noqa: E402
from oracle import stylegan2_oracle as so  # noqa: E402
from synthesis_in_style_b200 import _lib  # noqa: E402
from synthesis_in_style_b200 import dataset_gan as pg  # noqa: E402
from synthesis_in_style_b200.model import Generator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--ref-batch', type=int, default=2)
    ap.add_argument('--iters', type=int, default=5)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    size = 256
    spec = so.GeneratorSpec(size, 512, 8, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    g = Generator(size, 512, 8)
    g.load_state_dict(sd)
    g = g.to(dev).eval()
    torch.manual_seed(1)
    with torch.no_grad():
        _, acts = g([torch.randn(a.batch, 512).to(dev)], return_intermediate_activations=True, noise=[n.to(dev) for n in so.make_noise(spec)])
    F = sum(t.shape[1] for t in acts.values())
    states = [dg.init_classifier_state(F, 3, seed=50 + i, base_seed=49) for i in range(3)]
    ens = pg.PixelEnsembleClassifier(3, 0, 0)
    for st in states:
        net = pg.PixelClassifier(3, F)
        net.load_state_dict(st)
        ens.add_network(net.eval())
    for _ in range(2):
        labels, _, _ = ens.predict_label_images(acts, size)
    torch.cuda.synchronize()
    _lib.profile_enable(True); _lib.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        labels, _, _ = ens.predict_label_images(acts, size)
    e1.record(); torch.cuda.synchronize()
    prof = _lib.profile_collect(); _lib.profile_enable(False)
    ens.check(dev)
    ms = e0.elapsed_time(e1) / a.iters
    # the reference's algorithm on this GPU
    models = [dg.ClassifierParams({k: v.to(dev) for k, v in st.items()}) for st in states]
    sub = {k: v[:a.ref_batch] for k, v in acts.items()}
    with torch.no_grad():
        dg.predict_labels(models, sub, size)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        want, margin, _ = dg.predict_labels(models, sub, size)
        r1.record(); torch.cuda.synchronize()
    ref_ms = r0.elapsed_time(r1)
    safe = margin > 1e-3
    agree = float((labels[:a.ref_batch].float() == want).float().mean())
    exact = bool((labels[:a.ref_batch].float()[safe] == want[safe]).all())
    first_layer_flops = 2.0 * a.batch * 384 * sum(t.shape[1] * t.shape[-1] ** 2 for t in acts.values())
    print(json.dumps({'op': 'dataset_gan_label', 'image_size': size, 'batch': a.batch, 'features': F, 'networks': 3,
                      'ms_per_batch': round(ms, 3), 'images_per_s': round(a.batch / ms * 1e3, 1),
                      'ms_by_category': {k: round(v[0] / a.iters, 3) for k, v in prof.items() if v[1]},
                      'first_layer_alg_TFLOP/s': round(first_layer_flops / (prof['conv_tc'][0] / a.iters * 1e-3) / 1e12, 1),
                      'reference_algorithm_same_gpu': {'batch': a.ref_batch, 'ms_per_image': round(ref_ms / a.ref_batch, 2),
                                                       'images_per_s': round(a.ref_batch / ref_ms * 1e3, 1),
                                                       'note': 'materialised [B,S,S,F] features + three fp32 MLPs in PyTorch (TF32 off)'},
                      'labels_equal_where_margin_gt_1e-3': exact, 'label_agreement': round(agree, 6)}))


if __name__ == '__main__':
    main()
