"""CPU restatement of the reference's DatasetGAN labeller.  TEST INFRASTRUCTURE ONLY.

  scf/create_dataset_for_segmentation.py:28-49      get_dataset_gan_params: one bilinear nn.Upsample(scale_factor=S/size) per capture
  scf/data/dataset_gan_dataset.py:12-34             scale_activations: upsample every capture to S x S, concatenate channels -> [B,S,S,F]
  scf/networks/pixel_classifier/model.py:60-121     PixelClassifier: Linear(F,128) ReLU BN Linear(128,32) ReLU BN Linear(32,n)   (n < 32)
  scf/networks/base_segmenter.py:55-58              predict_classes: argmax of the (log-)softmax
  scf/networks/pixel_classifier/model.py:40-50      PixelEnsembleClassifier.predict_classes: torch.mode over the networks
  scf/segmentation/dataset_gan_segmenter.py:34-60   predict_labels, label_images_to_color_images, create_segmentation_image
Pinned by tests/golden/golden_dataset_gan_v1.npz (tests/golden/make_golden_dataset_gan.py runs the reference's own
classes on the CPU with their hard-coded .cuda() placements neutralised).
"""
from typing import Dict, List, Sequence, Tuple

import numpy
import torch
import torch.nn.functional as F


def scale_activations(activations: Dict[int, torch.Tensor], image_size: int) -> torch.Tensor:
    """dataset_gan_dataset.py:12-34 for one batch: [B, S, S, sum of channels], captures in dict order."""
    feats = []
    for act in activations.values():
        scale = image_size / act.shape[-1]
        up = F.interpolate(act, scale_factor=scale, mode='bilinear') if scale != 1 else F.interpolate(act, scale_factor=1.0, mode='bilinear')
        feats.append(torch.moveaxis(up, 1, -1))
    return torch.cat(feats, dim=-1)


class ClassifierParams:
    """Weights of one PixelClassifier in eval mode (state-dict keys `layers.{0,2,3,5,6}.*`)."""

    def __init__(self, sd: Dict[str, torch.Tensor]):
        g = lambda k: sd[k].detach().float()
        self.w1, self.b1 = g('layers.0.weight'), g('layers.0.bias')
        self.bn1 = (g('layers.2.weight'), g('layers.2.bias'), g('layers.2.running_mean'), g('layers.2.running_var'))
        self.w2, self.b2 = g('layers.3.weight'), g('layers.3.bias')
        self.bn2 = (g('layers.5.weight'), g('layers.5.bias'), g('layers.5.running_mean'), g('layers.5.running_var'))
        self.w3, self.b3 = g('layers.6.weight'), g('layers.6.bias')


def batch_norm_eval(x, bn, eps=1e-5):
    gamma, beta, mean, var = bn
    return F.batch_norm(x, mean, var, gamma, beta, False, 0.0, eps)


def classifier_logits(p: ClassifierParams, x: torch.Tensor) -> torch.Tensor:
    """PixelClassifier.forward, model.py:117-118 (eval-mode BatchNorm1d = affine with the running statistics)."""
    h = batch_norm_eval(F.relu(F.linear(x, p.w1, p.b1)), p.bn1)
    h = batch_norm_eval(F.relu(F.linear(h, p.w2, p.b2)), p.bn2)
    return F.linear(h, p.w3, p.b3)


def ensemble_votes(models: Sequence[ClassifierParams], x: torch.Tensor, chunk: int = 1 << 15):
    """Per-network argmax classes [N, n_models] (float, as the reference stores them) and the per-network top-2 logit
    margins, chunked over pixels."""
    votes, margins = [], []
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        v, m = [], []
        for p in models:
            logits = classifier_logits(p, xs)
            v.append(torch.max(F.softmax(logits, dim=1), dim=1)[1].float())
            top = logits.topk(min(2, logits.shape[1]), dim=1).values
            m.append(top[:, 0] - top[:, 1] if top.shape[1] > 1 else torch.full_like(top[:, 0], float('inf')))
        votes.append(torch.stack(v, dim=1))
        margins.append(torch.stack(m, dim=1))
    return torch.cat(votes), torch.cat(margins)


def predict_labels(models: Sequence[ClassifierParams], activations: Dict[int, torch.Tensor], image_size: int):
    """predict_labels, dataset_gan_segmenter.py:34-41: label image [B,S,S] (float class ids, torch.mode = the smallest of
    the most frequent votes) and the smallest top-2 margin over the networks (for the tolerance-aware comparison)."""
    feats = scale_activations(activations, image_size)
    b = feats.shape[0]
    votes, margins = ensemble_votes(models, feats.reshape(b * image_size * image_size, -1))
    labels = torch.mode(votes).values.reshape(b, image_size, image_size)
    return labels, margins.min(dim=1).values.reshape(b, image_size, image_size), votes.reshape(b, image_size, image_size, -1)


def label_images_to_color_images(label_images: torch.Tensor, class_to_color: Dict[str, Tuple[int, int, int]]) -> numpy.ndarray:
    """dataset_gan_segmenter.py:43-53 on [B,S,S] labels."""
    b, h, w = label_images.shape
    out = numpy.zeros((b, h, w, 3), dtype='uint8')
    out[:, :, :] = class_to_color['background']
    for class_id, (name, color) in enumerate(class_to_color.items()):
        if name == 'background':
            continue
        out[(label_images == class_id).numpy()] = color
    return out


def create_segmentation_image(models, activations, image_size, class_to_color):
    labels, _, _ = predict_labels(models, activations, image_size)
    return label_images_to_color_images(labels, class_to_color), []


def init_classifier_state(feature_size: int, n_class: int, seed: int, base_seed: int = None, jitter: float = 0.15) -> Dict[str, torch.Tensor]:
    """A synthetic trained-looking PixelClassifier state dict (Linear weights ~ N(0, 1/sqrt(fan_in)), non-trivial
    BatchNorm statistics) in the reference's key layout.  With `base_seed` the Linear weights are those of the
    base_seed network plus `jitter` of fresh noise, so that the networks of an ensemble mostly agree, as trained ones do."""
    def draw(s):
        g = torch.Generator().manual_seed(s)
        sd = {}
        dims = [(feature_size, 128), (128, 32), (32, n_class)]
        for idx, (fi, fo) in zip((0, 3, 6), dims):
            sd[f'layers.{idx}.weight'] = torch.randn(fo, fi, generator=g) / fi ** 0.5
            sd[f'layers.{idx}.bias'] = torch.randn(fo, generator=g) * 0.1
        for idx, f in ((2, 128), (5, 32)):
            sd[f'layers.{idx}.weight'] = 1 + 0.2 * torch.randn(f, generator=g)
            sd[f'layers.{idx}.bias'] = 0.1 * torch.randn(f, generator=g)
            sd[f'layers.{idx}.running_mean'] = 0.3 * torch.randn(f, generator=g)
            sd[f'layers.{idx}.running_var'] = 0.5 + torch.rand(f, generator=g)
            sd[f'layers.{idx}.num_batches_tracked'] = torch.tensor(100)
        return sd
    own = draw(seed)
    if base_seed is None:
        return own
    base = draw(base_seed)
    return {k: (base[k] + jitter * own[k] if k.endswith('.weight') and k.split('.')[1] in ('0', '3', '6') else base[k]) for k in base}
