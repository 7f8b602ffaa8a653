"""On-disk output: path rule, side-by-side PNGs, running ids across ranks (gloo, world size 2),
train / val JSON.  CPU only."""
import json
import os
import subprocess
import sys

import numpy
import pytest
from PIL import Image

from synthesis_in_style_b200 import dataset_writer as dw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLORS = {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'}


def test_path_rule_and_name_format(tmp_path):
    """create_dataset_for_segmentation.py:84-99: <base>/<id//100000>/<id//1000>/<zero-padded id>.png"""
    assert dw.name_format_for(100) == '{id:04d}.png' and dw.name_format_for(1000000) == '{id:07d}.png'
    p = dw.image_file_name(123456, tmp_path, dw.name_format_for(500000))
    assert p == tmp_path / '1' / '123' / '123456.png'
    assert dw.image_file_name(7, tmp_path, dw.name_format_for(100)) == tmp_path / '0' / '0' / '0007.png'


def test_side_by_side_png_round_trip(tmp_path):
    rng = numpy.random.RandomState(0)
    gen = rng.randint(0, 256, size=(3, 16, 16, 3), dtype=numpy.uint8)
    lab = rng.randint(0, 256, size=(3, 16, 16, 3), dtype=numpy.uint8)
    files = dw.save_generated_images(gen, lab, 998, tmp_path, 2000)
    assert [f.relative_to(tmp_path).as_posix() for f in files] == ['0/0/0998.png', '0/0/0999.png', '0/1/1000.png']
    for i, f in enumerate(files):
        data = numpy.array(Image.open(f))
        assert data.shape == (16, 32, 3)
        assert numpy.array_equal(data[:, :16], gen[i]) and numpy.array_equal(data[:, 16:], lab[i])


def test_assign_round_ids_replays_the_sequential_loop():
    # 2 ranks, target 10: round 0 keeps 4 + 3, round 1 keeps 4 (reaches 11 >= 10) so rank 1's batch never existed
    starts, n, done = dw.assign_round_ids([4, 3], 0, 10)
    assert (starts, n, done) == ([0, 4], 7, False)
    starts, n, done = dw.assign_round_ids([4, 2], 7, 10)
    assert (starts, n, done) == ([7, None], 11, True)
    # a batch that keeps nothing still advances the round
    assert dw.assign_round_ids([0, 5], 3, 100) == ([3, 3], 8, False)


def _batches(seed, n_batches, batch, size=8):
    rng = numpy.random.RandomState(seed)
    out = []
    for _ in range(n_batches):
        gen = rng.randint(0, 256, size=(batch, size, size, 3), dtype=numpy.uint8)
        lab = rng.randint(0, 256, size=(batch, size, size, 3), dtype=numpy.uint8)
        drop = sorted(set(rng.randint(0, batch, size=rng.randint(0, 3)).tolist()))
        out.append((gen, lab, drop))
    return out


def _tree(root):
    return {p.relative_to(root).as_posix(): p.read_bytes() for p in sorted(root.glob('**/*.png'))}


def test_two_ranks_write_the_single_process_dataset(tmp_path):
    """Batch index sharding + the kept-count all-gather reproduce the reference's single-stream ids and files."""
    single = tmp_path / 'single'
    w = dw.DatasetWriter(single, 21)
    for gen, lab, drop in _batches(5, 12, 4):
        w.add(gen, lab, drop)
        if w.finished:
            break
    w.flush()
    script = tmp_path / 'w.py'
    script.write_text(
        "import os, sys, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
        "from synthesis_in_style_b200 import dataset_writer as dw\n"
        "from test_dataset_writer import _batches\n"
        "dist.init_process_group('gloo')\n"
        "r, W = dist.get_rank(), dist.get_world_size()\n"
        f"w = dw.DatasetWriter({str(tmp_path / 'multi')!r}, 21, r, W)\n"
        "for i, (gen, lab, drop) in enumerate(_batches(5, 12, 4)):\n"
        "    if i % W != r: continue\n"
        "    w.add(gen, lab, drop)\n"
        "    if w.finished: break\n"
        "w.flush(); dist.barrier(); dist.destroy_process_group(); print('ok', r, w.n)\n")
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                          '--master-port', '29741', str(script)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    a, b = _tree(single), _tree(tmp_path / 'multi')
    assert len(a) >= 21 and a.keys() == b.keys()
    assert all(a[k] == b[k] for k in a)


def test_has_class_flags_and_split(tmp_path):
    size = 12
    def pair(label):
        return numpy.concatenate([numpy.zeros((size, size, 3), numpy.uint8), label], axis=1)
    blank = numpy.zeros((size, size, 3), numpy.uint8)
    one_px = blank.copy(); one_px[3, 3] = (0, 0, 255)                 # a single pixel: 1-point contour, not counted
    block = blank.copy(); block[2:6, 2:7] = (0, 0, 255); block[8:10, 1:3] = (255, 0, 0)
    assert dw.determine_classes_in_image(pair(blank), COLORS) == {'has_printed_text': False, 'has_handwritten_text': False}
    assert dw.determine_classes_in_image(pair(one_px), COLORS) == {'has_printed_text': False, 'has_handwritten_text': False}
    assert dw.determine_classes_in_image(pair(block), COLORS) == {'has_printed_text': True, 'has_handwritten_text': True}
    for i in range(20):
        dw.save_image(pair(block if i % 2 else blank), i, tmp_path, dw.name_format_for(20))
    train, val = dw.write_train_val_split(tmp_path, COLORS, seed=1)
    t, v = json.load(open(train)), json.load(open(val))
    assert len(t) == 18 and len(v) == 2
    assert {e['file_name'] for e in t + v} == {f'0/0/{i:04d}.png' for i in range(20)}
    assert all(set(e) == {'file_name', 'has_printed_text', 'has_handwritten_text'} for e in t + v)
    by_name = {e['file_name']: e for e in t + v}
    assert by_name['0/0/0001.png']['has_printed_text'] and not by_name['0/0/0000.png']['has_printed_text']


def test_png_encoders_decode_equal(tmp_path):
    """Every PNG writer (the reference's PIL call, libpng's fast setting, this module's own zlib level 1 / stored
    writers) produces files that decode to the same pixels; RGB, grey and RGBA."""
    from PIL import Image
    rng = numpy.random.RandomState(3)
    image = rng.randint(0, 256, size=(33, 65, 3)).astype(numpy.uint8)
    image[5:20] = image[4]                                   # repeated rows: the 'Up' filter path
    gray = rng.randint(0, 256, size=(16, 17)).astype(numpy.uint8)
    rgba = rng.randint(0, 256, size=(9, 8, 4)).astype(numpy.uint8)
    for enc in ('pil', 'cv2', 'fast', 'stored'):
        for i, arr in enumerate((image, gray, rgba)):
            f = dw.save_image(arr, i, tmp_path / enc, '{id}.png', encoder=enc)
            assert numpy.array_equal(numpy.array(Image.open(f)), arr), (enc, arr.shape)
    assert dw.PNG_ENCODER in ('fast', 'cv2', 'stored', 'pil')
    with pytest.raises(ValueError):
        dw.save_image(image, 9, tmp_path, '{id}.png', encoder='bmp')
    with pytest.raises(ValueError):
        dw.png_bytes(image.astype(numpy.float32))


def test_native_writer_equals_python_writer(tmp_path):
    """sis_png_write_pairs (native threads, no GIL) writes the files save_generated_images writes: same names, same
    pixels; rows select the kept images; stored and deflated variants; an unwritable path is reported."""
    rng = numpy.random.RandomState(5)
    gen = rng.randint(0, 256, size=(5, 24, 20, 3), dtype=numpy.uint8)
    lab = numpy.zeros((5, 24, 28, 3), dtype=numpy.uint8)
    lab[:, 4:9, 3:17] = (0, 0, 255)
    rows = [0, 2, 3]
    for level, sub in ((1, 'fast'), (0, 'stored')):
        files = dw.save_generated_images_native(gen, lab, rows, 998, tmp_path / sub, 2000, level=level, n_threads=3)
        assert [f.relative_to(tmp_path / sub).as_posix() for f in files] == ['0/0/0998.png', '0/0/0999.png', '0/1/1000.png']
        for f, r in zip(files, rows):
            data = numpy.array(Image.open(f))
            assert data.shape == (24, 48, 3)
            assert numpy.array_equal(data[:, :20], gen[r]) and numpy.array_equal(data[:, 20:], lab[r])
    assert dw.save_generated_images_native(gen, lab, [], 0, tmp_path / 'none', 10) == []
    blocker = tmp_path / 'blocked'
    blocker.mkdir()
    (blocker / '0').write_text('a file where a directory is needed')
    with pytest.raises((RuntimeError, OSError)):
        dw.save_generated_images_native(gen, lab, rows, 0, blocker, 10)
    # the writer object on both paths
    for native in (True, False):
        w = dw.DatasetWriter(tmp_path / f'w{int(native)}', 8, native=native)
        w.add(gen, lab[:, :, :20], [1])
        w.add(gen, lab[:, :, :20], [])
        w.flush()
        assert w.n == 9 and w.finished and w.files_written == 9
    a, b = _tree(tmp_path / 'w1'), _tree(tmp_path / 'w0')
    assert a.keys() == b.keys() and len(a) == 9
    for k in a:
        assert numpy.array_equal(numpy.array(Image.open(tmp_path / 'w1' / k)), numpy.array(Image.open(tmp_path / 'w0' / k)))
