"""DatasetGAN labeller: host mirror of the reference's other `segmenter_type`.

  scf/networks/pixel_classifier/model.py:13-121      PixelClassifier, PixelEnsembleClassifier
  scf/segmentation/dataset_gan_segmenter.py:12-60    DatasetGANSegmenter
  scf/create_dataset_for_segmentation.py:28-49       get_dataset_gan_params (feature_size, upsamplers)
The device work is one C-ABI call (`sis_pixel_ensemble_label`, csrc/dataset_gan.cu): the first Linear of every network
runs at each capture's native resolution on the tcgen05 GEMM, a tail kernel does the bilinear gather, the rest of the
MLPs, the argmax and the mode vote.  The upsampled [B, S, S, F] feature tensor of the reference is never built.
Inference only (the networks are used in eval mode, as DatasetGANSegmenter.load_ensemble sets them).
"""
import ctypes
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy
import torch
from torch import nn

from . import _lib
from .labelling import BaseDatasetSegmenter

PARAM_KEYS = ('layers.0.weight', 'layers.0.bias', 'layers.2.weight', 'layers.2.bias', 'layers.2.running_mean', 'layers.2.running_var',
              'layers.3.weight', 'layers.3.bias', 'layers.5.weight', 'layers.5.bias', 'layers.5.running_mean', 'layers.5.running_var',
              'layers.6.weight', 'layers.6.bias')


class PixelClassifier(nn.Module):
    """Parameter container with the reference's layout (model.py:60-86, the numpy_class < 32 variant), so that
    `load_state_dict(checkpoint['network_i'])` works.  It holds weights; the arithmetic happens in the ensemble call."""

    def __init__(self, numpy_class: int, dim: int):
        super().__init__()
        if numpy_class >= 32:
            raise NotImplementedError('the 256/128 classifier for >= 32 classes is not on this path')
        self.numpy_class, self.dim = numpy_class, dim
        self.layers = nn.Sequential(nn.Linear(dim, 128), nn.ReLU(), nn.BatchNorm1d(num_features=128), nn.Linear(128, 32), nn.ReLU(),
                                    nn.BatchNorm1d(num_features=32), nn.Linear(32, numpy_class))

    def init_weights(self, init_type: str = 'normal', gain: float = 0.02):
        """model.py:88-115 for Linear layers."""
        for m in self.layers:
            if isinstance(m, nn.Linear):
                {'normal': lambda w: nn.init.normal_(w, 0.0, gain), 'xavier': lambda w: nn.init.xavier_normal_(w, gain=gain),
                 'kaiming': lambda w: nn.init.kaiming_normal_(w, a=0, mode='fan_in'),
                 'orthogonal': lambda w: nn.init.orthogonal_(w, gain=gain)}[init_type](m.weight.data)
                nn.init.constant_(m.bias.data, 0.0)

    def forward(self, x):
        raise RuntimeError('PixelClassifier is evaluated through PixelEnsembleClassifier.predict_label_images (B200 path); '
                           'there is no PyTorch / CPU fallback')


class PixelEnsembleClassifier:
    """model.py:13-50.  `predict_label_images` replaces predict_classes over the materialised feature tensor."""

    def __init__(self, numpy_class: int, dim: int, number_of_models: int = 0):
        self.numpy_class, self.dim = numpy_class, dim
        self.networks: Dict[str, PixelClassifier] = {}
        self.last_net_id = 0
        for i in range(number_of_models):
            net = PixelClassifier(numpy_class, dim)
            net.init_weights()
            self.networks[f'network_{i}'] = net
            self.last_net_id += 1
        self._handle = None
        self._feature_size = None
        self._built_from = None

    def get_networks(self):
        return self.networks

    def set_network(self, network_name: str, network: PixelClassifier):
        self.networks[network_name] = network
        self._release()

    def add_network(self, network: PixelClassifier):
        self.last_net_id += 1
        self.networks[f'network_{self.last_net_id}'] = network
        self._release()

    def _release(self):
        if self._handle is not None:
            _lib.load().sis_pixel_ensemble_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _signature(self, feature_size: int):
        """Identity of everything the native ensemble was built from: a replaced tensor (`load_state_dict` on a fresh
        module, `set_network`) changes `data_ptr`, an in-place update (`load_state_dict`, an optimiser step, `copy_`) bumps
        `_version`."""
        sig = [feature_size, self.numpy_class]
        for name, net in self.networks.items():
            sd = net.state_dict()
            sig.append((name, tuple((key, sd[key].data_ptr(), sd[key]._version, tuple(sd[key].shape)) for key in PARAM_KEYS)))
        return tuple(sig)

    def _sync(self, feature_size: int, stream):
        signature = self._signature(feature_size)
        if self._handle is not None and self._built_from == signature:
            return
        self._release()
        lib = _lib.load()
        nets = list(self.networks.values())
        if not nets:
            raise RuntimeError('the ensemble holds no networks')
        handle = ctypes.c_void_p()
        _lib.check(lib.sis_pixel_ensemble_create(ctypes.byref(handle), len(nets), feature_size, self.numpy_class))
        for i, net in enumerate(nets):
            sd = net.state_dict()
            for key in PARAM_KEYS:
                t = sd[key].detach().to('cpu', torch.float32).contiguous()
                _lib.check(lib.sis_pixel_ensemble_set_param(handle, i, key.encode(), ctypes.c_void_p(t.data_ptr()), t.numel()))
        _lib.check(lib.sis_pixel_ensemble_prepare(handle, stream))
        self._handle, self._feature_size, self._built_from = handle, feature_size, signature

    def predict_label_images(self, activations: Dict[int, torch.Tensor], image_size: int, colors: Optional[Sequence[Tuple[int, int, int]]] = None,
                             want_votes: bool = False):
        """(labels uint8 [B,S,S], votes uint8 [B,S,S,n] or None, colour images uint8 [B,S,S,3] or None), on the device."""
        acts = [activations[k] for k in activations]          # dict order, as scale_activations iterates
        for t in acts:
            _lib.require_cuda(t, 'activation')
        acts = [t.contiguous().float() for t in acts]
        dev = acts[0].device
        batch = acts[0].shape[0]
        n = len(acts)
        stream = _lib.current_stream_ptr(dev)
        with torch.cuda.device(dev):
            self._sync(sum(t.shape[1] for t in acts), stream)
            labels = torch.empty(batch, image_size, image_size, dtype=torch.uint8, device=dev)
            votes = torch.empty(batch, image_size, image_size, len(self.networks), dtype=torch.uint8, device=dev) if want_votes else None
            color_images = torch.empty(batch, image_size, image_size, 3, dtype=torch.uint8, device=dev) if colors is not None else None
            host_colors = numpy.ascontiguousarray(numpy.asarray(colors, dtype=numpy.uint8).reshape(-1)) if colors is not None else None
            ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in acts])
            chans = (ctypes.c_int * n)(*[t.shape[1] for t in acts])
            ress = (ctypes.c_int * n)(*[t.shape[-1] for t in acts])
            _lib.check(_lib.load().sis_pixel_ensemble_label(
                self._handle, n, ptrs, chans, ress, batch, image_size, _lib.ptr(labels), _lib.ptr(votes),
                ctypes.c_void_p(host_colors.ctypes.data) if host_colors is not None else ctypes.c_void_p(0), _lib.ptr(color_images), stream))
        return labels, votes, color_images

    def check(self, device=None):
        if self._handle is not None:
            _lib.check(_lib.load().sis_pixel_ensemble_check(self._handle, _lib.current_stream_ptr(device)))


def get_dataset_gan_params(activations: Dict[int, torch.Tensor], image_size: int) -> Dict:
    """create_dataset_for_segmentation.py:28-49: feature size and one bilinear upsampler per capture (kept for config
    compatibility; the B200 path folds the upsampling into its tail kernel and never applies these modules)."""
    return {'feature_size': sum(a.shape[1] for a in activations.values()),
            'upsamplers': [nn.Upsample(scale_factor=image_size / a.shape[-1], mode='bilinear') for a in activations.values()]}


class DatasetGANSegmenter(BaseDatasetSegmenter):
    """dataset_gan_segmenter.py:12-60."""

    def __init__(self, base_dir, image_size: int, class_to_color_map: Dict, classifier_path: Optional[str] = None,
                 feature_size: Optional[int] = None, upsamplers: Optional[List[nn.Upsample]] = None,
                 ensemble: Optional[PixelEnsembleClassifier] = None):
        super().__init__(base_dir, image_size, class_to_color_map)
        self.upsamplers = upsamplers
        self.ensemble = ensemble if ensemble is not None else self.load_ensemble(classifier_path, feature_size)

    def load_ensemble(self, path: str, feature_size: int) -> PixelEnsembleClassifier:
        """:22-32: every checkpoint entry whose key contains 'network' (and not 'optimizer') is one classifier."""
        n_class = len(self.class_to_color_map)
        ensemble = PixelEnsembleClassifier(n_class, self.image_size, 0)
        checkpoint = torch.load(path, map_location='cpu')
        for key in checkpoint.keys():
            if 'network' in key and 'optimizer' not in key:
                model = PixelClassifier(n_class, feature_size)
                model.load_state_dict(checkpoint[key])
                model.eval()
                ensemble.add_network(model)
        return ensemble

    @torch.no_grad()
    def predict_labels(self, activations: Dict[int, torch.Tensor]) -> torch.Tensor:
        """:34-41, from the captures instead of the scaled feature tensor: label images [B,S,S] (uint8)."""
        return self.ensemble.predict_label_images(activations, self.image_size)[0]

    def label_images_to_color_images(self, label_images: torch.Tensor) -> numpy.ndarray:
        """:43-53 (host version, for label images that did not come with colours)."""
        labels = label_images.detach().cpu().numpy().reshape(label_images.shape[0], self.image_size, self.image_size)
        table = numpy.zeros((max(len(self.class_to_color_map), int(labels.max()) + 1), 3), dtype=numpy.uint8)
        for class_id, color in enumerate(self.class_to_color_map.values()):
            table[class_id] = color
        return table[labels]

    @torch.no_grad()
    def create_segmentation_image(self, activations: Dict[int, torch.Tensor]):
        """:55-60: (uint8 [B,S,S,3] colour label images, []) -- this segmenter never drops an image."""
        colors = list(self.class_to_color_map.values())
        _, _, color_images = self.ensemble.predict_label_images(activations, self.image_size, colors=colors)
        return color_images.cpu().numpy(), []
