"""CPU: host-side logic of the package (no kernels run): module tree / init order, sharding, catalog ingestion,
the world_size-2 statistics reduction over gloo, and the loud failure without CUDA."""
import copy
import json
import os
import pickle
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from synthesis_in_style_b200 import dataset_creation as dc
from synthesis_in_style_b200 import labelling
from synthesis_in_style_b200.model import Generator
from synthesis_in_style_b200.op import fused_leaky_relu, upfirdn2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_init_matches_reference_weights(golden):
    torch.manual_seed(0)
    g = Generator(16, 64, 2, channel_multiplier=2)
    sd = g.state_dict()
    for key in ('style.1.weight', 'conv1.conv.weight', 'convs.2.conv.weight', 'noises.noise_2'):
        idx = golden[f'g16trunc/w/{key}/idx']
        np.testing.assert_array_equal(sd[key].reshape(-1).numpy()[idx], golden[f'g16trunc/w/{key}/val'])
    assert (g.size, g.style_dim, g.log_size, g.num_layers, g.n_latent) == (16, 64, 4, 5, 6)
    assert g.activation_shape(0) == (512, 4) and g.activation_shape(5) == (512, 16)


def test_state_dict_keys_match_reference_layout():
    from oracle import stylegan2_oracle as so
    g = Generator(32, 64, 2)
    ref = so.init_state_dict(so.GeneratorSpec(32, 64, 2, 2), seed=0)
    assert set(g.state_dict().keys()) == set(ref.keys())
    for k, v in g.state_dict().items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
    g.load_state_dict(ref)          # g_ema-style load
    g2 = copy.deepcopy(g)           # the native plan is not shared by copies
    assert g2._plan is not g._plan and g2._plan.handle is None


def test_no_cpu_fallback():
    g = Generator(8, 32, 1)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match='CUDA tensor'):
            g([torch.randn(1, 32)])
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        fused_leaky_relu(torch.randn(2, 3), torch.zeros(3))
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        upfirdn2d(torch.randn(1, 1, 4, 4), torch.ones(2, 2))
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        labelling.FactorCatalog(2, np.eye(2, 4, dtype=np.float32)).predict(torch.randn(1, 4, 2, 2))


def test_latent_stream_semantics_and_sharding():
    g = Generator(8, 32, 1)
    cfg = {'batch_size': 3, 'latent_size': 32}
    single = []
    it = iter(dc.build_latent_and_noise_generator(g, cfg, seed=1))
    for _ in range(6):
        single.append(next(it))
    # consecutive CPU randn(B, 512) calls are positional: the reference stream
    torch.manual_seed(1)
    z0 = torch.randn(3, 32)
    assert torch.equal(single[0].latent, z0)
    assert [tuple(n.shape) for n in single[0].noise] == [(1, 1, 4, 4), (1, 1, 8, 8), (1, 1, 8, 8)]
    # a wrapper object with .decoder is accepted like the reference's StyleganAutoencoder
    it2 = iter(dc.build_latent_and_noise_generator(types.SimpleNamespace(decoder=g), cfg, seed=1))
    assert torch.equal(next(it2).latent, z0)
    # rank shards replay the same stream and partition the batch indices
    got = {}
    for rank in range(2):
        s = dc.sharded_latent_stream(g, cfg, 1, rank, 2)
        for _ in range(3):
            idx, b = next(s)
            got[idx] = b
    assert sorted(got) == list(range(6))
    for i in range(6):
        assert torch.equal(got[i].latent, single[i].latent)
        for a, b in zip(got[i].noise, single[i].noise):
            assert torch.equal(a, b)
    assert dc.owned_batches(1, 4, 10) == [1, 5, 9]
    assert sorted(sum((dc.owned_batches(r, 4, 10) for r in range(4)), [])) == list(range(10))


def test_latents_dataclass():
    lat = dc.Latents(torch.randn(2, 4), [torch.randn(2, 1, 4, 4)])
    one = lat[1]
    assert one.latent.shape == (1, 4) and one.noise[0].shape == (1, 1, 4, 4)
    assert isinstance(lat.numpy().latent, np.ndarray)


class _FakeKMeans:
    def __init__(self, c):
        self.cluster_centers_ = c


class _FakeCatalog:
    def __init__(self, c):
        self._factorization = _FakeKMeans(c)
        self.annotations = {}


def test_catalog_pickle_is_read_without_sklearn(tmp_path):
    # a stand-in for catalogs/{k}.pkl (create_semantic_segmentation.py:123-137): classes that do not exist at load time
    mod = types.ModuleType('gone_module')
    _FakeKMeans.__module__ = 'gone_module'
    _FakeCatalog.__module__ = 'gone_module'
    mod._FakeKMeans, mod._FakeCatalog = _FakeKMeans, _FakeCatalog
    sys.modules['gone_module'] = mod
    c8 = np.random.RandomState(0).randn(4, 16).astype(np.float32)
    c9 = np.random.RandomState(1).randn(4, 16).astype(np.float32)
    try:
        blob = pickle.dumps({'8': _FakeCatalog(c8), '9': _FakeCatalog(c9), 'id_to_size_map': {'8': 64}})
    finally:
        del sys.modules['gone_module']
    (tmp_path / 'catalogs').mkdir()
    (tmp_path / 'catalogs' / '4.pkl').write_bytes(blob)
    cents = labelling.extract_centroids_from_pickle(tmp_path / 'catalogs' / '4.pkl')
    assert sorted(cents) == ['8', '9']
    np.testing.assert_array_equal(cents['8'], c8)
    (tmp_path / 'merged_classes_4.json').write_text(json.dumps(
        {'8': {'0': 'background', '1': 'printed_text', '2': 'handwritten_text', '3': 'background'},
         '9': {'0': 'background', '1': 'printed_text', '2': 'printed_text', '3': 'background'}}))
    seg = labelling.ClusterSegmenter(tmp_path, 64, {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'},
                                     keys_for_class_determination=['8'], keys_for_finegrained_segmentation=['9'],
                                     num_clusters=4, keys_to_merge={})
    assert sorted(seg.catalog) == ['8', '9'] and seg.catalog['9'].k == 4
    assert dict(seg.class_label_map['8']) == {'background': [0, 3], 'printed_text': [1], 'handwritten_text': [2]}
    assert seg.class_id_map == {'background': 0, 'printed_text': 1, 'handwritten_text': 2}
    assert seg.class_to_color_map['printed_text'] == (0, 0, 255)
    # an unlabelled class name is rejected like the reference's sanity assert
    (tmp_path / 'merged_classes_4.json').write_text(json.dumps({'8': {'0': 'x'}, '9': {'0': 'background'}}))
    with pytest.raises(AssertionError):
        labelling.ClusterSegmenter(tmp_path, 64, {'background': '#000000'}, ['8'], ['9'], 4)


def test_stats_vector_layout():
    sizes = {'8': 3, '12': 2}
    vec = torch.tensor([5, 6, 1, 2, 3, 64, 2])     # sorted keys: '12' then '8'
    s = dc.split_stats(vec, sizes)
    assert s == {'cluster_pixels': {'12': [5, 6], '8': [1, 2, 3]}, 'pairs': 64, 'batches': 2}


def test_stats_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from synthesis_in_style_b200 import dataset_creation as dc\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank()\n"
        "owned = dc.owned_batches(r, dist.get_world_size(), 7)\n"
        "vec = torch.tensor([10 * (r + 1), r, len(owned) * 4, len(owned)], dtype=torch.int64)\n"
        "dc.reduce_stats(vec)\n"
        "assert vec.tolist() == [30, 1, 28, 7], vec.tolist()\n"
        "dist.barrier(); dist.destroy_process_group()\n"
        "print('ok', r)\n")
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                          '--master-port', '29731', str(script)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count('ok') == 2
