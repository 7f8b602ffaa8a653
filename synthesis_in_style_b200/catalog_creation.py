"""Catalog creation: spherical mini-batch k-means over captured activations, on the device.

Mirrors the surface of
  scf/segmentation/gan_local_edit/spherical_kmeans.py:159-312   MiniBatchSphericalKMeans.fit
  scf/segmentation/gan_local_edit/factor_catalog.py:14-68       FactorCatalog(k).fit_predict(X, raw=True)
  scf/create_semantic_segmentation.py:114-137                   find_and_render_clusters, save_catalogs
The reference's `fit` is sklearn 0.24.2's MiniBatchKMeans loop (k-means++ init on a random subset, per-centre learning
rate 1 / count, random reassignment of starved centres, EWA-inertia early stopping) run on unit-normalised rows with the
centres re-normalised after every step.  The same algorithm runs here with its data on the GPU: the assign step of
every mini-batch, of the validation set and of the final labelling is `sis_label_assign` (the hot path's labelling
kernel: centres are unit vectors, so the nearest centre by squared distance is the spherical assignment), the centre
update is a segmented sum.  PARITY UNPINNED: the reference's result depends on sklearn 0.24.2 internals and numpy's
RandomState stream (neither is available here: sklearn is 1.9, `_k_means_fast` is gone), so only properties are
tested -- unit-norm centres, every label is the nearest centre, planted clusters are recovered, the inertia does not
exceed that of the initialisation.  The `.npz` this writes is what `labelling.load_catalog_file` reads.
"""
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy
import torch

from . import _lib
from .labelling import label_assign


def _normalize_rows(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def _assign(points_t: torch.Tensor, centers: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """`points_t` [C, n] (channel-major, as the labelling kernel reads NCHW) -> (labels int64 [n], squared distance to
    the nearest centre fp32 [n])."""
    c, n = points_t.shape
    pad = (-n) % 4
    if pad:
        points_t = torch.cat([points_t, points_t[:, :pad]], dim=1)
    act = points_t.reshape(1, c, points_t.shape[1], 1)
    ids, out = label_assign(act, centers, want_ids_i64=True, want_margin=False)
    labels = ids.reshape(-1)[:n]
    pts = points_t[:, :n]
    d = (pts * pts).sum(0) - 2.0 * (pts * centers.t()[:, labels]).sum(0) + (centers * centers).sum(1)[labels]
    return labels, d.clamp_min(0)


class MiniBatchSphericalKMeans:
    """sklearn-style estimator with the reference's attributes (`cluster_centers_`, `labels_`, `inertia_`, `n_iter_`)."""

    def __init__(self, n_clusters: int = 8, random_state: Optional[int] = 0, batch_size: int = 100, max_iter: int = 100, n_init: int = 3,
                 init_size: Optional[int] = None, reassignment_ratio: float = 0.01, max_no_improvement: int = 10, compute_labels: bool = True,
                 **_unused):
        self.n_clusters, self.random_state, self.batch_size, self.max_iter, self.n_init = n_clusters, random_state, batch_size, max_iter, n_init
        self.init_size, self.reassignment_ratio, self.max_no_improvement = init_size, reassignment_ratio, max_no_improvement
        self.compute_labels = compute_labels
        self.cluster_centers_ = None
        self.labels_ = None

    def _kmeans_pp(self, x: torch.Tensor, gen: torch.Generator) -> torch.Tensor:
        """k-means++ seeding on the rows of `x` (unit vectors): D^2 sampling with the usual 2 + log k local trials."""
        n, k = x.shape[0], self.n_clusters
        trials = 2 + int(numpy.log(k))
        first = int(torch.randint(n, (1,), generator=gen, device=x.device))
        centers = [x[first]]
        closest = ((x - centers[0]) ** 2).sum(1)
        for _ in range(1, k):
            cand = torch.multinomial(closest.clamp_min(1e-30), trials, replacement=True, generator=gen)
            d_cand = ((x[None, :, :] - x[cand][:, None, :]) ** 2).sum(2)          # [trials, n]
            pot = torch.minimum(closest[None], d_cand).sum(1)
            best = int(pot.argmin())
            centers.append(x[cand[best]])
            closest = torch.minimum(closest, d_cand[best])
        return torch.stack(centers)

    def _step(self, xb: torch.Tensor, centers: torch.Tensor, counts: torch.Tensor, reassign: bool, gen: torch.Generator) -> float:
        """One mini-batch update (sklearn's `_mini_batch_step`, dense variant), centres modified in place."""
        labels, dist = _assign(xb.t().contiguous(), centers)
        k = centers.shape[0]
        if reassign and self.reassignment_ratio > 0:
            starved = counts < self.reassignment_ratio * counts.max()
            if int(starved.sum()) > 0.5 * xb.shape[0]:
                keep = torch.argsort(counts)[int(0.5 * xb.shape[0]):]
                starved[keep] = False
            n_re = int(starved.sum())
            if n_re:
                pick = torch.randperm(xb.shape[0], generator=gen, device=xb.device)[:n_re]
                centers[starved] = xb[pick]
                counts[starved] = counts[~starved].min() if bool((~starved).any()) else 0
        sums = torch.zeros_like(centers).index_add_(0, labels, xb)
        n_in = torch.bincount(labels, minlength=k).to(centers.dtype)
        hit = n_in > 0
        centers[hit] = centers[hit] * counts[hit, None] + sums[hit]
        counts += n_in
        centers[hit] = centers[hit] / counts[hit, None]
        return float(dist.sum())

    def fit(self, X, y=None, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError('sample weights are not used by the catalog creation path')
        _lib.require_cuda(X, 'X')
        x = _normalize_rows(X.float().contiguous())
        n, k = x.shape[0], self.n_clusters
        if n < k:
            raise ValueError(f'n_samples={n} should be >= n_clusters={k}')
        gen = torch.Generator(device=x.device)
        gen.manual_seed(int(self.random_state or 0))
        init_size = min(n, self.init_size or 3 * self.batch_size)
        valid = x[torch.randint(n, (init_size,), generator=gen, device=x.device)]
        valid_t = valid.t().contiguous()
        best = None
        for _ in range(self.n_init):
            sub = x[torch.randint(n, (init_size,), generator=gen, device=x.device)]
            centers = _normalize_rows(self._kmeans_pp(sub, gen))
            counts = torch.zeros(k, device=x.device)
            self._step(valid, centers, counts, False, gen)
            centers = _normalize_rows(centers)
            inertia = float(_assign(valid_t, centers)[1].sum())
            if best is None or inertia < best[0]:
                best = (inertia, centers.clone(), counts.clone())
        self.init_inertia_, centers, counts = best
        n_batches = int(numpy.ceil(n / self.batch_size))
        n_iter = int(self.max_iter * n_batches)
        ewa, ewa_min, no_improvement = None, None, 0
        alpha = min(1.0, self.batch_size * 2.0 / (n + 1))
        it = 0
        for it in range(n_iter):
            idx = torch.randint(n, (self.batch_size,), generator=gen, device=x.device)
            centers = _normalize_rows(centers)
            reassign = (it + 1) % (10 + int(counts.min())) == 0
            batch_inertia = self._step(x[idx], centers, counts, reassign, gen) / self.batch_size
            centers = _normalize_rows(centers)
            ewa = batch_inertia if ewa is None else ewa * (1 - alpha) + batch_inertia * alpha
            if ewa_min is None or ewa < ewa_min:
                ewa_min, no_improvement = ewa, 0
            else:
                no_improvement += 1
            if self.max_no_improvement is not None and no_improvement >= self.max_no_improvement:
                break
        self.n_iter_ = it + 1
        self._centers_device = centers
        self.cluster_centers_ = centers.cpu().numpy()
        if self.compute_labels:
            labels, dist = _assign(x.t().contiguous(), centers)
            self.labels_, self.inertia_ = labels.cpu().numpy(), float(dist.sum())
        return self

    def predict(self, X) -> numpy.ndarray:
        x = _normalize_rows(X.float().contiguous())
        return _assign(x.t().contiguous(), self._centers_device.to(x.device))[0].cpu().numpy()


def fit_catalog(activations: torch.Tensor, num_clusters: int, random_state: int = 0, **kmeans_args):
    """`FactorCatalog(k).fit_predict(X, raw=True)` for one layer: ([N, k, H, W] one-hot heat maps, centres float32 [k, C])."""
    n, c, h, w = activations.shape
    flat = activations.permute(0, 2, 3, 1).reshape(-1, c)          # ptutils.partial_flat
    km = MiniBatchSphericalKMeans(num_clusters, random_state=random_state, compute_labels=True, **kmeans_args).fit(flat)
    labels = torch.from_numpy(km.labels_).to(activations.device).reshape(n, h, w)
    heat = torch.nn.functional.one_hot(labels, num_clusters).permute(0, 3, 1, 2).float()
    return heat, km.cluster_centers_


def find_clusters(all_activations: Dict[int, torch.Tensor], num_clusters: int, min_size: int = 0, **kmeans_args):
    """create_semantic_segmentation.py:96-137 without the rendering: ({layer: heat maps}, {str(layer): centres} plus
    'id_to_size_map').  Layers whose maps are not larger than `min_size` are skipped (`strip_activations`, :93-94)."""
    heat, catalogs, sizes = {}, {}, {}
    for key, act in all_activations.items():
        if min_size and not (act.shape[-2] > min_size and act.shape[-1] > min_size):
            continue
        heat[key], catalogs[str(key)] = fit_catalog(act, num_clusters, **kmeans_args)
        sizes[key] = f'{act.shape[-2]}x{act.shape[-1]}'
    return heat, catalogs, sizes


def save_catalogs(catalogs: Dict[str, numpy.ndarray], num_clusters: int, dest_dir) -> Path:
    """`<dest_dir>/<k>.npz` ({layer: float32 [k, C]}): the catalog format `ClusterSegmenter.load_catalog` prefers (the
    reference pickles sklearn-backed FactorCatalog objects, create_semantic_segmentation.py:131-137)."""
    dest_dir = Path(dest_dir)
    dest_dir.mkdir(parents=True, exist_ok=True)
    path = dest_dir / f'{num_clusters}.npz'
    numpy.savez(path, **{k: numpy.asarray(v, dtype=numpy.float32) for k, v in catalogs.items()})
    return path
