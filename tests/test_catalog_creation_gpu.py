"""Catalog creation on the GPU.  The reference's k-means depends on sklearn 0.24.2 internals and
numpy's RandomState stream, neither available here: parity is unpinned and the tests check properties."""
import numpy as np
import pytest
import torch

from oracle import labelling_oracle as lo
from synthesis_in_style_b200 import catalog_creation as cc
from synthesis_in_style_b200 import labelling

pytestmark = pytest.mark.gpu


def planted(n_per, k, c, device, noise=0.05, seed=0):
    g = torch.Generator().manual_seed(seed)
    centers = torch.nn.functional.normalize(torch.randn(k, c, generator=g), dim=1)
    x = centers.repeat_interleave(n_per, 0) + noise * torch.randn(k * n_per, c, generator=g)
    scale = 0.5 + torch.rand(k * n_per, 1, generator=g) * 3          # spherical: the norm must not matter
    perm = torch.randperm(k * n_per, generator=g)
    return (x * scale)[perm].to(device), torch.arange(k).repeat_interleave(n_per)[perm], centers


def test_planted_clusters_are_recovered(cuda_device):
    x, truth, centers = planted(400, 6, 48, cuda_device)
    km = cc.MiniBatchSphericalKMeans(6, random_state=0).fit(x)
    cent = torch.from_numpy(km.cluster_centers_)
    torch.testing.assert_close(cent.norm(dim=1), torch.ones(6), rtol=0, atol=1e-5)          # unit-norm centres
    # every found centre sits on one planted direction, and each planted direction is found once
    sim = cent @ centers.t()
    match = sim.argmax(1)
    assert sorted(match.tolist()) == list(range(6)) and float(sim.max(1).values.min()) > 0.99
    # labels = nearest centre of the normalised points (the oracle's argmin), consistent with the planted assignment
    xn = torch.nn.functional.normalize(x.cpu(), dim=1)
    want = lo.pairwise_distances(xn, cent).argmin(1).numpy()
    assert (km.labels_ == want).mean() >= 0.999
    assert (match[torch.from_numpy(km.labels_)] == truth).float().mean() >= 0.99
    # both sit on the noise floor of the planted data; the fit must not have drifted away from the seeding's quality
    assert km.inertia_ / len(x) <= 1.25 * km.init_inertia_ / min(len(x), 300)
    assert np.array_equal(km.predict(x[:50]), km.labels_[:50])


def test_catalog_round_trip_into_the_segmenter(cuda_device, tmp_path):
    """fit on generator-like activations -> <ssd>/catalogs/<k>.npz -> ClusterSegmenter labels with those centres."""
    torch.manual_seed(0)
    acts = {4: torch.randn(3, 32, 16, 16, device=cuda_device), 5: torch.randn(3, 32, 16, 16, device=cuda_device), 0: torch.randn(3, 32, 4, 4, device=cuda_device)}
    heat, catalogs, sizes = cc.find_clusters(acts, 4, min_size=4, random_state=0)
    assert sorted(catalogs) == ['4', '5'] and sizes == {4: '16x16', 5: '16x16'}               # the 4x4 map is stripped
    assert heat[4].shape == (3, 4, 16, 16) and float(heat[4].sum(1).min()) == 1.0
    path = cc.save_catalogs(catalogs, 4, tmp_path / 'catalogs')
    assert path.name == '4.npz'
    names = {'background': '#000000', 'printed_text': '#0000FF'}
    (tmp_path / 'merged_classes_4.json').write_text(
        '{"4": {"0": "background", "1": "printed_text", "2": "background", "3": "printed_text"},'
        ' "5": {"0": "background", "1": "printed_text", "2": "background", "3": "printed_text"}}')
    seg = labelling.ClusterSegmenter(tmp_path, 16, names, ['4'], ['5'], 4)
    for layer in ('4', '5'):
        ids = seg.catalog[layer].predict(acts[int(layer)])
        # raw activations against unit centres (the hot path's rule) agree with the k-means labels of the normalised
        # points wherever the norm does not flip the argmin: on random data that is the large majority
        want = heat[int(layer)].argmax(1)
        assert (ids == want).float().mean() > 0.5
        np.testing.assert_allclose(np.linalg.norm(catalogs[layer], axis=1), 1.0, atol=1e-5)
