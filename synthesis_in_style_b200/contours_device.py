"""Contour stage on the device: the masks never leave the GPU.

Host mirror of `sis_contour_stage` (csrc/contours.cu, include/sis_b200.h).  Same inputs and results as
`contours.segment_masks` (the polygon implementation that mirrors
  scf/segmentation/black_white_handwritten_printed_text_segmenter.py:77-99
  scf/segmentation/base_cluster_based_dataset_segmenter.py:148-450),
i.e. (uint8 [B,S,S,3] colour label images, ids of the images to drop), computed from label maps instead of cv2 polygons.

The device does not order contours, and the reference's drop rule reads the FIRST contour of a class.  Whenever that order
could matter for an image (one contour of a class exceeds 95 % of the image in both directions and another one does
not), or a capacity of the device stage is exceeded, the kernel flags the image and `segment` sends just that image
through `contours.segment_masks`, so the results are the reference's in every case.  `run` never synchronises: the merge
fixpoint is controlled on the device.
"""
import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy
import torch

from . import _lib
from .contours import ContourConfig

FLAG_KEEP, FLAG_DROP, FLAG_HOST = 0, 1, 2


class DeviceContourStage:
    """Workspace + argument marshalling for one (batch size, image size, config)."""

    def __init__(self, cfg: ContourConfig, fine_class: str = 'printed_text'):
        self.cfg = cfg
        self.classes = [name for name in cfg.class_to_color_map if name != 'background']
        if fine_class not in self.classes:
            raise KeyError(f'fine-grained class {fine_class!r} is not in class_to_color_map')
        self.fine_class = fine_class
        colors = [cfg.class_to_color_map['background']] + [cfg.class_to_color_map[n] for n in self.classes]
        self._colors = bytes(int(v) for c in colors for v in c)
        self._workspace = None
        self._info = None

    def supports(self, class_names: Dict[str, Sequence[str]]) -> bool:
        """The device stage needs every class under every key it reads (the reference tolerates a key without a class:
        its contour lists are then shorter; that case takes the host path)."""
        keys = list(self.cfg.keys_for_class_determination) + list(self.cfg.keys_for_finegrained_segmentation)
        return all(key in class_names and all(n in class_names[key] for n in self.classes) for key in keys)

    def run(self, stacked: Dict[str, Tuple[Sequence[str], torch.Tensor]]):
        """stacked: {key: (class names, uint8 [n_class, B, S, S] on the device)} (the labelling kernels' own layout).
        Returns (uint8 [B,S,S,3] label images, int32 [B] flags), both on the device, enqueued on the current stream."""
        cfg = self.cfg
        det_keys, fine_keys = list(cfg.keys_for_class_determination), list(cfg.keys_for_finegrained_segmentation)
        first = stacked[det_keys[0]][1]
        _lib.require_cuda(first, 'masks')
        device = first.device
        B, S = first.shape[1], first.shape[-1]
        if S != cfg.image_size or first.shape[-2] != S:
            raise RuntimeError(f'masks must be at image size {cfg.image_size} (got {tuple(first.shape[-2:])})')

        def plane(key, name):
            names, masks = stacked[key]
            _lib.require_cuda(masks, 'masks')
            if masks.dtype not in (torch.uint8, torch.bool) or masks.shape[1:] != (B, S, S):
                raise RuntimeError(f'masks of key {key}: expected uint8 [n_class, {B}, {S}, {S}]')
            m = masks[list(names).index(name)]
            m = m.view(torch.uint8) if m.dtype == torch.bool else m
            return m if m.is_contiguous() else m.contiguous()

        det = [plane(k, n) for k in det_keys for n in self.classes]
        fine = [plane(k, self.fine_class) for k in fine_keys]
        n_cls = len(self.classes)
        last_names = list(stacked[fine_keys[-1]][0])
        order = [n for n in last_names if n != 'background']          # render_segmentation_image iterates this dict (:421)
        rank = (ctypes.c_int * n_cls)(*[order.index(n) for n in self.classes])
        lib = _lib.load()
        need = ctypes.c_int64(0)
        _lib.check(lib.sis_contour_stage_workspace_bytes(B, S, n_cls, len(det_keys), len(fine_keys), ctypes.byref(need)))
        if self._workspace is None or self._workspace.numel() < need.value or self._workspace.device != device:
            self._workspace = torch.empty(need.value, dtype=torch.uint8, device=device)
        out = torch.empty(B, S, S, 3, dtype=torch.uint8, device=device)
        flags = torch.empty(B, dtype=torch.int32, device=device)
        info = torch.zeros(3, dtype=torch.int32, device=device)
        det_p = (ctypes.c_void_p * len(det))(*[_lib.ptr(t) for t in det])
        fine_p = (ctypes.c_void_p * len(fine))(*[_lib.ptr(t) for t in fine])
        with torch.cuda.device(device):
            _lib.check(lib.sis_contour_stage(det_p, fine_p, B, S, n_cls, len(det_keys), len(fine_keys),
                                             self.classes.index(self.fine_class), int(bool(cfg.only_keep_overlapping)),
                                             float(cfg.min_class_contour_area), self._colors, rank, _lib.ptr(self._workspace),
                                             self._workspace.numel(), _lib.ptr(out), _lib.ptr(flags), _lib.ptr(info),
                                             _lib.current_stream_ptr(device)))
        self._info = info
        return out, flags

    @property
    def last_info(self):
        """(shapes found, fixpoint rounds that did work, reason the batch went to the host path or 0) of the last `run`;
        reading it waits for that call's kernels."""
        return (0, 0, 0) if self._info is None else tuple(int(v) for v in self._info.cpu())


def _stack(predicted_clusters) -> Dict[str, Tuple[List[str], torch.Tensor]]:
    """PredictedClusters {key: {class: bool/uint8 [B,S,S]}} -> the stacked layout."""
    out = {}
    for key, per_class in predicted_clusters.items():
        names = list(per_class)
        planes = [(m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)) for m in per_class.values()]
        out[key] = (names, torch.stack(planes, dim=0))
    return out


def warm_worker() -> int:
    """Nothing but the imports a fall-back task needs (run once per pool worker before the pipeline starts)."""
    import os
    import time

    from . import contours  # noqa: F401
    time.sleep(0.05)          # keep this worker busy so the next warm-up task lands on another one
    return os.getpid()


def host_fallback(stacked_host: Dict[str, Tuple[Sequence[str], numpy.ndarray]], image_ids: Sequence[int], cfg: ContourConfig):
    """`contours.segment_masks` for single images; stacked_host holds uint8 [n_class, B, S, S] numpy arrays.
    Returns {image id: (uint8 [S,S,3], dropped)}."""
    from . import contours
    keys = set(cfg.keys_for_class_determination) | set(cfg.keys_for_finegrained_segmentation)
    out = {}
    for b in image_ids:
        per_image = {key: {name: masks[j, b:b + 1] for j, name in enumerate(names)}
                     for key, (names, masks) in stacked_host.items() if key in keys}
        image, drop = contours.segment_masks(per_image, 1, cfg)
        out[b] = (image[0], bool(drop))
    return out


def segment(stacked_or_predicted, batch_size: int, cfg: ContourConfig, stage: Optional[DeviceContourStage] = None):
    """Drop-in for `contours.segment_masks` on device masks: (uint8 [B,S,S,3] numpy, sorted ids of the images to drop).
    Accepts the stacked layout or a PredictedClusters dict of device tensors."""
    first = next(iter(stacked_or_predicted.values()))
    stacked = stacked_or_predicted if isinstance(first, tuple) else _stack(stacked_or_predicted)
    stage = stage or DeviceContourStage(cfg)
    names = {key: list(n) for key, (n, _) in stacked.items()}
    if not stage.supports(names):
        host = {key: (n, m.cpu().numpy()) for key, (n, m) in stacked.items()}
        res = host_fallback(host, range(batch_size), cfg)
        return numpy.stack([res[b][0] for b in range(batch_size)]), [b for b in range(batch_size) if res[b][1]]
    images_d, flags_d = stage.run(stacked)
    images, flags = images_d.cpu().numpy(), flags_d.cpu().numpy()
    drop = [int(b) for b in numpy.flatnonzero(flags == FLAG_DROP)]
    undecided = [int(b) for b in numpy.flatnonzero(flags == FLAG_HOST)]
    if undecided:
        keys = set(cfg.keys_for_class_determination) | set(cfg.keys_for_finegrained_segmentation)
        idx = torch.as_tensor(undecided, device=images_d.device)
        host = {key: (n, m.index_select(1, idx).cpu().numpy()) for key, (n, m) in stacked.items() if key in keys}
        res = host_fallback(host, range(len(undecided)), cfg)
        for j, b in enumerate(undecided):
            images[b] = res[j][0]
            if res[j][1]:
                drop.append(b)
    return images, sorted(drop)
