"""CPU oracle: activation -> semantic-class labelling.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch-CPU restatement of
  scf/segmentation/gan_local_edit/factor_catalog.py:47-75  (FactorCatalog.predict / pairwise_distance)
  scf/segmentation/gan_local_edit/ptutils.py:25-28          (partial_flat)
  scf/segmentation/base_cluster_based_dataset_segmenter.py:56-67,119-138 (class map inversion, predict_clusters)
  scf/segmentation/base_dataset_segmenter.py:32-42          (resize_to_image_size, nearest)
  scf/segmentation/black_white_handwritten_printed_text_segmenter.py:31-40 (merge_sub_images)
  scf/create_dataset_for_segmentation.py:39-44 + scf/data/dataset_gan_dataset.py:12-34
      (bilinear feature upsample of the DatasetGAN variant; the north star's "bilinear upsample" mode)
Pinned by tests/golden/make_golden.py against those functions run from /root/reference.
"""
from collections import defaultdict
from functools import reduce
from typing import Dict, List

import torch
import torch.nn.functional as F


def partial_flat(x: torch.Tensor) -> torch.Tensor:
    """ptutils.py:25-28: NCHW -> [N*H*W, C]."""
    return x.permute(0, 2, 3, 1).contiguous().view(-1, x.shape[1])


def pairwise_distances(flat: torch.Tensor, centroids: torch.Tensor, chunk: int = 1 << 16) -> torch.Tensor:
    """factor_catalog.py:47-58: ((A[:,None]-B[None])**2).sum(-1), chunked over rows to bound the
    [N,k,C] temporary (chunking does not change any row's arithmetic)."""
    outs = []
    b = centroids.unsqueeze(0)
    for s in range(0, flat.shape[0], chunk):
        a = flat[s:s + chunk].unsqueeze(1)
        outs.append(((a - b) ** 2.0).sum(dim=-1))
    return torch.cat(outs, 0) if outs else flat.new_zeros((0, centroids.shape[0]))


def predict(x: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """FactorCatalog.predict, factor_catalog.py:69-75: argmin (ties -> first) reshaped to [B,H,W] int64."""
    b, _, h, w = x.shape
    d = pairwise_distances(partial_flat(x), centroids)
    return torch.argmin(d, dim=1).reshape(b, h, w)


def predict_with_margin(x: torch.Tensor, centroids: torch.Tensor):
    """ids plus the margin d2-d1 used by the north-star label criterion."""
    b, _, h, w = x.shape
    d = pairwise_distances(partial_flat(x), centroids)
    ids = torch.argmin(d, dim=1)
    if d.shape[1] > 1:
        srt = d.sort(1).values
        margin = srt[:, 1] - srt[:, 0]
    else:
        margin = torch.full((d.shape[0],), float('inf'))
    return ids.reshape(b, h, w), margin.reshape(b, h, w)


def invert_class_label_map(class_label_map: Dict[str, Dict[str, str]]) -> Dict[str, Dict[str, List[int]]]:
    """load_class_label_map, base_cluster_based_dataset_segmenter.py:56-67:
    {layer: {cluster_id_str: class_name}} -> {layer: {class_name: [cluster ids]}} (first-seen order)."""
    inverted = {}
    for key, sub in class_label_map.items():
        inv = defaultdict(list)
        for sub_key, label_name in sub.items():
            inv[label_name].append(int(sub_key))
        inverted[key] = inv
    return inverted


def predict_clusters(activations: Dict[int, torch.Tensor], catalog: Dict[str, torch.Tensor],
                     class_label_map: Dict[str, Dict[str, List[int]]]) -> Dict[str, Dict[str, torch.Tensor]]:
    """predict_clusters, base_cluster_based_dataset_segmenter.py:119-138.
    `catalog` maps layer -> centroid matrix [k,C] (the only part of FactorCatalog used at inference)."""
    out = {}
    acts = {str(k): v for k, v in activations.items()}
    for layer_id, centroids in catalog.items():
        membership = predict(acts[layer_id], centroids)
        per_class = {}
        for class_name, class_ids in class_label_map[layer_id].items():
            masks = [membership == cid for cid in class_ids]
            per_class[class_name] = reduce(torch.bitwise_or, masks, torch.zeros_like(masks[0], dtype=torch.bool))
        out[layer_id] = per_class
    return out


def resize_to_image_size(tensors, image_size: int):
    """resize_to_image_size, base_dataset_segmenter.py:32-42 (F.interpolate default = nearest)."""
    resized = {}
    for key, class_tensors in tensors.items():
        r = {}
        for class_name, t in class_tensors.items():
            if t.shape[-1] < image_size:
                t = F.interpolate(t[:, None, ...].type(torch.uint8), (image_size, image_size)).type(t.dtype).squeeze(1)
            r[class_name] = t
        resized[key] = r
    return resized


def merge_sub_images(predicted, keys_to_merge: Dict[str, List[str]], class_names: List[str]):
    """merge_sub_images, black_white_handwritten_printed_text_segmenter.py:31-40."""
    for dst, keys in keys_to_merge.items():
        subs = [predicted[k] for k in keys]
        merged = {}
        for cn in class_names:
            ts = [s[cn] for s in subs]
            merged[cn] = reduce(torch.bitwise_or, ts[1:], ts[0])
        predicted[dst] = merged
    return predicted


def prepare_image_segmentation(activations, catalog, class_label_map, image_size: int):
    """prepare_image_segmentation, base_cluster_based_dataset_segmenter.py:140-146."""
    return resize_to_image_size(predict_clusters(activations, catalog, class_label_map), image_size)


def bilinear_then_predict(x: torch.Tensor, centroids: torch.Tensor, image_size: int) -> torch.Tensor:
    """North-star / DatasetGAN-style feature path: nn.Upsample(scale_factor=S/size, mode='bilinear')
    (align_corners=False; create_dataset_for_segmentation.py:39-44, data/dataset_gan_dataset.py:12-34)
    followed by the same nearest-centroid argmin."""
    if x.shape[-1] != image_size:
        x = F.interpolate(x, scale_factor=image_size / x.shape[-1], mode='bilinear')
    return predict(x, centroids)


def make_image(x: torch.Tensor) -> torch.Tensor:
    """`pytorch_training.images.make_image` (un-vendored, unpinned dependency; call site
    create_dataset_for_segmentation.py:135). Restated from the public repo's behaviour:
    clamp(-1,1) -> (x+1)/2 -> *255 -> uint8 (truncate) -> NHWC.  uint8 image parity is UNPINNED
    (SURVEY.md §8c iii): compare float images; treat uint8 as +-1 LSB."""
    x = torch.clamp(x, -1, 1)
    x = (x + 1) / 2
    x = x * 255
    return x.type(torch.uint8).permute(0, 2, 3, 1).contiguous()
