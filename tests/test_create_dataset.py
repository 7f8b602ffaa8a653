"""The creation entry point (create_dataset_for_segmentation.py's surface): CPU tests of the flags, paths, config
loading and the dispatch on `segmenter_type` with the reference's creation JSON (a byte copy of
configs/dataset_creation/stylegan2_cluster_based_bw_hwp_wpi.json under tests/golden/); one GPU run end to end."""
import json
import os

import numpy as np
import pytest
import torch

from synthesis_in_style_b200 import create_dataset as cd
from synthesis_in_style_b200 import labelling

HERE = os.path.dirname(os.path.abspath(__file__))
CREATION_JSON = os.path.join(HERE, 'golden', 'creation_config_stylegan2_cluster_based_bw_hwp_wpi.json')


def write_catalog(base, layers, channels, k, seed=5):
    gen = torch.Generator().manual_seed(seed)
    (base / 'catalogs').mkdir(parents=True)
    np.savez(base / 'catalogs' / f'{k}.npz', **{l: torch.nn.functional.normalize(torch.randn(k, channels[l], generator=gen), dim=1).numpy()
                                               for l in layers})
    names = ['background', 'printed_text', 'handwritten_text']
    (base / f'merged_classes_{k}.json').write_text(json.dumps({l: {str(i): names[i % 3] for i in range(k)} for l in layers}))


def test_flags_and_defaults_are_the_references():
    p = cd.build_arg_parser()
    a = p.parse_args(['run/checkpoints/100.pt', 'creation.json'])
    assert (a.num_images, a.batch_size, a.device, a.num_clusters, a.truncate, a.debug) == (100, 10, 'cuda', -1, False, False)
    assert a.original_config_path is None and a.save_to is None and a.semantic_segmentation_base_dir is None
    a = p.parse_args(['c.pt', 'creation.json', '-op', 'orig.json', '-n', '7', '-s', 'out', '-b', '4', '-d', '1', '--truncate',
                      '--only-create-train-val-split', '--num-clusters', '20', '--classifier-path', 'cls.pt', '-ssd', 'sem'])
    assert (a.num_images, a.batch_size, a.device, a.num_clusters, a.truncate, a.only_create_train_val_split) == (7, 4, '1', 20, True, True)
    assert str(a.original_config_path) == 'orig.json' and str(a.semantic_segmentation_base_dir) == 'sem' and a.classifier_path == 'cls.pt'


def test_base_dirs_and_config_loading(tmp_path):
    run = tmp_path / 'run'
    (run / 'checkpoints').mkdir(parents=True)
    (run / 'config').mkdir()
    (run / 'config' / 'config.json').write_text(json.dumps({'image_size': 256, 'latent_size': 512}))
    (run / 'config' / 'args.json').write_text(json.dumps({'stylegan_variant': 2, 'latent_size': 256}))
    p = cd.build_arg_parser()
    a = p.parse_args([str(run / 'checkpoints' / '100.pt'), 'creation.json'])
    images, sem = cd.get_base_dirs(a)
    assert images == run / 'generated_images' and images.is_dir() and sem == run / 'semantic_segmentation'
    a = p.parse_args([str(run / 'checkpoints' / '100.pt'), 'creation.json', '-ssd', str(tmp_path / 'x' / 'sem'), '-s', str(tmp_path / 'out')])
    images, sem = cd.get_base_dirs(a)
    assert images == tmp_path / 'out' and sem == tmp_path / 'x' / 'sem'
    assert cd.load_config(str(run / 'checkpoints' / '100.pt')) == {'image_size': 256, 'latent_size': 256, 'stylegan_variant': 2}
    alt = tmp_path / 'orig.yaml'
    alt.write_text('image_size: 128\nlatent_size: 512\n')
    assert cd.load_config(None, alt) == {'image_size': 128, 'latent_size': 512}
    with pytest.raises(RuntimeError):
        cd.load_config(None, None)
    with pytest.raises(FileNotFoundError):
        cd.load_config(str(tmp_path / 'nowhere' / 'checkpoints' / '1.pt'))


def test_segmenter_dispatch_reads_the_reference_creation_json(tmp_path):
    with open(CREATION_JSON) as f:
        creation = json.load(f)
    assert creation['segmenter_type'] == 'black_white_handwritten_printed' and creation['keys_for_class_determination'] == ['8', '9']
    channels = {'8': 512, '9': 512, '12': 128, '13': 128}
    write_catalog(tmp_path, list(channels), channels, 20)
    args = cd.build_arg_parser().parse_args(['c.pt', CREATION_JSON, '--num-clusters', '20'])
    seg = cd.get_dataset_segmenter(args, creation, 256, tmp_path)
    assert isinstance(seg, labelling.ClusterSegmenter)
    assert seg.keys_for_class_determination == ['8', '9'] and seg.keys_for_finegrained_segmentation == ['12', '13']
    assert seg.min_class_contour_area == 50 and seg.only_keep_overlapping is False and seg.num_clusters == 20
    assert sorted(seg.catalog) == ['12', '13', '8', '9'] and seg.catalog['12'].k == 20
    assert list(seg.class_id_map) == ['background', 'printed_text', 'handwritten_text']
    with pytest.raises(NotImplementedError):
        cd.get_dataset_segmenter(args, dict(creation, segmenter_type='something_else'), 256, tmp_path)
    del creation['only_keep_overlapping']
    with pytest.raises(AssertionError):
        cd.get_dataset_segmenter(args, creation, 256, tmp_path)


@pytest.mark.gpu
def test_end_to_end_small_dataset(cuda_device, tmp_path):
    """`main` on a 32^2 random-init generator: PNG tree with the reference's names, train/val split, coco_gt.json."""
    from PIL import Image
    creation = {'class_to_color_map': {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'},
                'keys_for_finegrained_segmentation': ['6', '7'], 'keys_for_class_determination': ['4', '5'], 'keys_to_merge': {},
                'segmenter_type': 'black_white_handwritten_printed', 'only_keep_overlapping': False, 'min_class_contour_area': 2, 'seed': 1}
    cfg_path = tmp_path / 'creation.json'
    cfg_path.write_text(json.dumps(creation))
    orig = tmp_path / 'orig.json'
    orig.write_text(json.dumps({'image_size': 32, 'latent_size': 512, 'stylegan_variant': 2}))
    sem = tmp_path / 'run' / 'semantic_segmentation'
    write_catalog(sem, ['4', '5', '6', '7'], {l: 512 for l in '4567'}, 4)
    out = tmp_path / 'out'
    args = cd.build_arg_parser().parse_args(['random-init:0', str(cfg_path), '-op', str(orig), '-n', '24', '-b', '4', '-s', str(out),
                                             '--num-clusters', '4', '-ssd', str(sem), '--truncate'])
    stats = cd.main(args)
    files = sorted(out.glob('**/*.png'))
    assert stats['images_kept_all_ranks'] >= 24 and len(files) == stats['images_kept_all_ranks']
    assert files[0].relative_to(out).as_posix() == '0/0/0000.png'
    assert np.array(Image.open(files[0])).shape == (32, 64, 3)
    train, val = json.loads((out / 'train.json').read_text()), json.loads((out / 'val.json').read_text())
    assert len(train) == int(len(files) * 0.9) and len(train) + len(val) == len(files)
    assert set(train[0]) == {'file_name', 'has_printed_text', 'has_handwritten_text'}
    coco = json.loads((out / 'coco_gt.json').read_text())
    assert len(coco['images']) == len(val) and [c['name'] for c in coco['categories']] == list(creation['class_to_color_map'])
    for ann in coco['annotations']:
        assert ann['area'] >= 0 and len(ann['bbox']) == 4 and ann['segmentation']['size'] == [32, 32]
    # --only-create-train-val-split re-derives the same split from the files on disk
    args2 = cd.build_arg_parser().parse_args(['random-init:0', str(cfg_path), '-op', str(orig), '-s', str(out), '-ssd', str(sem),
                                              '--only-create-train-val-split'])
    cd.main(args2)
    assert json.loads((out / 'train.json').read_text()) == train
