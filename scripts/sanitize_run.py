#!/usr/bin/env python
"""Small end-to-end pass for `compute-sanitizer --tool memcheck`: one 256^2 generator forward at batch 1 with in-forward
labelling (every conv GEMM variant, blur, ToRGB, label kernels) and one DatasetGAN labelling call."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import dataset_gan_oracle as dg  # noqa: E402  (synthetic classifier weights only)
from synthesis_in_style_b200 import dataset_gan as pg, labelling  # noqa: E402
from synthesis_in_style_b200.model import Generator  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    spec, sd = bench.oracle_state()
    g = Generator(bench.SIZE, bench.STYLE_DIM, bench.N_MLP)
    g.load_state_dict(sd)
    g = g.to(dev).eval()
    catalog = {k: labelling.FactorCatalog(v.shape[0], v) for k, v in bench.synthetic_catalog().items()}
    seg = labelling.ClusterSegmenter(None, bench.SIZE, bench.COLORS, keys_for_class_determination=['8', '9'],
                                     keys_for_finegrained_segmentation=['12', '13'], num_clusters=4, keys_to_merge={},
                                     catalog=catalog, class_label_map={layer: bench.CLASS_MAP for layer in bench.LABEL_LAYERS})
    for batch in (1, 3):
        jobs = seg.make_label_jobs(g, batch)
        with torch.no_grad():
            img, acts = g([torch.randn(batch, 512, device=dev)], noise=g.make_noise(), return_intermediate_activations=True, label_jobs=jobs)
        torch.cuda.synchronize()
        print('forward ok', batch, tuple(img.shape), float(img.abs().max()))
    feat = sum(t.shape[1] for t in acts.values())
    ens = pg.PixelEnsembleClassifier(3, 0, 0)
    for i in range(3):
        net = pg.PixelClassifier(3, feat)
        net.load_state_dict(dg.init_classifier_state(feat, 3, seed=50 + i, base_seed=49))
        ens.add_network(net.eval())
    labels, _, _ = ens.predict_label_images(acts, bench.SIZE)
    ens.check(dev)
    torch.cuda.synchronize()
    print('dataset_gan ok', tuple(labels.shape), labels.unique().tolist())


if __name__ == '__main__':
    main()
