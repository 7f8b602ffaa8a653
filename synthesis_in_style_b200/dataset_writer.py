"""On-disk output of the dataset-creation loop: side-by-side PNGs, directory sharding,
running ids, train / val JSON.

Mirrors scf/create_dataset_for_segmentation.py
  :84-90    save_image                 <base>/<id // 100000>/<id // 1000>/<name_format>
  :93-99    save_generated_images      image || label concatenated along the width, id = batch_id + row
  :109-148  build_dataset loop         drop, running id = number of images kept so far, stop once >= num_images
  :151-206  create_dataset_json_data, main: shuffle under random.seed(config['seed']), 90 / 10 split, train.json / val.json
and scf/segmentation/evaluation/coco_gt.py:35-65 (`determine_classes_in_image`: a class is present when its colour has
an external contour with at least 3 points in the right half of the PNG).
coco_gt.json (coco_gt.py:67-134) is written by synthesis_in_style_b200/coco_gt.py.

PNG encoding is CPU-bound zlib work; `DatasetWriter` hands the rows of a batch to a thread pool (PIL releases the GIL
while it compresses), so files are written while the GPU and the contour workers produce the next batch.
Multi-GPU runs assign the reference's GLOBAL running ids: ranks exchange the kept count of their batch once per round
(`exchange_kept_counts`, an all-gather of one int64) and `assign_round_ids` replays the reference's sequential loop.
"""
import json
import random
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import cv2
import numpy
from PIL import Image

from .labelling import BaseDatasetSegmenter


def image_file_name(image_id: int, base_dir: Path, name_format: str = '{id}.png') -> Path:
    """save_image's path rule (:85-87)."""
    return Path(base_dir) / str(image_id // 100000) / str(image_id // 1000) / name_format.format(id=image_id)


def name_format_for(num_images: int) -> str:
    """save_generated_images' name format (:99): zero-padded to max(4, digits of num_images)."""
    return f'{{id:0{max(4, len(str(num_images)))}d}}.png'


# PNG encoder of save_image.  All of them write files that decode to the same pixels:
#   'pil'    the reference's call (Image.fromarray(image).save(dest), :88): zlib level 6 behind PIL's adaptive filter, 44 ms
#            for a 256 x 512 side-by-side image on one core, and PIL holds the GIL while it deflates, so a thread pool
#            scales 2x at best; 87 KB.  Use it to reproduce the reference's bytes.
#   'cv2'    libpng at its fast setting: 5-7 ms, GIL released, 105 KB.
#   'fast'   (default) this module's writer: 'Up' filter on every row (one vectorised subtraction), zlib level 1 with the
#            run-length strategy, the three chunks assembled by hand: 3.3 ms, 104 KB; zlib and numpy release the GIL, a thread pool scales with the cores.
#   'stored' the same writer at zlib level 0 (stored deflate blocks): 0.6 ms, 394 KB -- when the writer must keep up with
#            several GPUs on few host cores.
PNG_ENCODER = 'fast'
_PNG_COLOR_TYPE = {1: 0, 2: 4, 3: 2, 4: 6}          # channels -> PNG colour type (grey, grey+alpha, RGB, RGBA)


def png_bytes(image: numpy.ndarray, level: int = 1) -> bytes:
    """A complete PNG file (8-bit, non-interlaced) for an [H, W] or [H, W, C] uint8 array."""
    import struct
    import zlib
    if image.dtype != numpy.uint8 or image.ndim not in (2, 3):
        raise ValueError('png_bytes expects a uint8 [H, W] or [H, W, C] array')
    h, w = image.shape[:2]
    c = 1 if image.ndim == 2 else image.shape[2]
    if c not in _PNG_COLOR_TYPE or h == 0 or w == 0:
        raise ValueError(f'cannot write a PNG of shape {image.shape}')
    flat = numpy.ascontiguousarray(image).reshape(h, w * c)
    raw = numpy.empty((h, 1 + w * c), dtype=numpy.uint8)
    if level == 0:
        raw[:, 0] = 0                                     # filter 'None': nothing to gain without compression
        raw[:, 1:] = flat
    else:
        raw[:, 0] = 2                                     # filter 'Up': row minus the row above (mod 256); the first row's
        raw[0, 1:] = flat[0]                              # predecessor is all zeros
        numpy.subtract(flat[1:], flat[:-1], out=raw[1:, 1:])
    if level == 0:
        data = zlib.compress(raw, 0)
    else:                                                 # run-length strategy: faster and smaller than the default on 'Up' rows
        comp = zlib.compressobj(level, zlib.DEFLATED, 15, 9, zlib.Z_RLE)
        data = comp.compress(raw) + comp.flush()

    def chunk(tag: bytes, body: bytes) -> bytes:
        return struct.pack('>I', len(body)) + tag + body + struct.pack('>I', zlib.crc32(body, zlib.crc32(tag)))

    return b''.join((b'\x89PNG\r\n\x1a\n', chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, _PNG_COLOR_TYPE[c], 0, 0, 0)),
                     chunk(b'IDAT', data), chunk(b'IEND', b'')))


def encode_png(image: numpy.ndarray, dest, encoder: Optional[str] = None):
    encoder = encoder or PNG_ENCODER
    if encoder == 'pil':
        Image.fromarray(image).save(str(dest))
    elif encoder == 'cv2':
        data = image if image.ndim == 2 else cv2.cvtColor(image, cv2.COLOR_RGB2BGR if image.shape[2] == 3 else cv2.COLOR_RGBA2BGRA)
        if not cv2.imwrite(str(dest), data):
            raise OSError(f'could not write {dest}')
    elif encoder in ('fast', 'stored'):
        with open(str(dest), 'wb') as f:
            f.write(png_bytes(image, 1 if encoder == 'fast' else 0))
    else:
        raise ValueError(f'unknown PNG encoder {encoder!r}')


def save_image(image: numpy.ndarray, image_id: int, base_dir: Path, name_format: str = '{id}.png', encoder: Optional[str] = None) -> Path:
    dest = image_file_name(image_id, base_dir, name_format)
    dest.parent.mkdir(exist_ok=True, parents=True)
    encode_png(image, dest, encoder)
    return dest


def save_generated_images(generated_images: numpy.ndarray, semantic_segmentation_images: numpy.ndarray, batch_id: int,
                          base_dir: Path, num_images: int, pool=None) -> List:
    """:93-99.  With `pool` (a concurrent.futures executor) returns the futures of the per-file writes."""
    images = numpy.concatenate([generated_images, semantic_segmentation_images], axis=2)
    fmt = name_format_for(num_images)
    if pool is None:
        return [save_image(image, batch_id + idx, base_dir, fmt) for idx, image in enumerate(images)]
    return [pool.submit(save_image, image, batch_id + idx, base_dir, fmt) for idx, image in enumerate(images)]


def save_generated_images_native(generated_images: numpy.ndarray, semantic_segmentation_images: numpy.ndarray, rows: Sequence[int],
                                 batch_id: int, base_dir: Path, num_images: int, level: int = 1, n_threads: int = 4) -> List[Path]:
    """save_generated_images (:93-99) for the rows `rows` of a batch, through the library's native writer
    (`sis_png_write_pairs`, csrc/png_writer.cu): concatenation, filtering, deflate and the file writes run on native
    threads without the interpreter lock.  File i gets id batch_id + i.  Blocks until the files are on disk."""
    import ctypes

    from . import _lib
    gen = numpy.ascontiguousarray(generated_images, dtype=numpy.uint8)
    lab = numpy.ascontiguousarray(semantic_segmentation_images, dtype=numpy.uint8)
    if gen.ndim != 4 or lab.ndim != 4 or gen.shape[0] != lab.shape[0] or gen.shape[1] != lab.shape[1] or gen.shape[3] != lab.shape[3]:
        raise ValueError(f'expected two [B, H, W, C] batches of equal height and channels, got {gen.shape} and {lab.shape}')
    fmt = name_format_for(num_images)
    paths = [image_file_name(batch_id + i, base_dir, fmt) for i in range(len(rows))]
    for parent in {p.parent for p in paths}:
        parent.mkdir(exist_ok=True, parents=True)
    if not paths:
        return paths
    c_rows = (ctypes.c_int32 * len(rows))(*[int(r) for r in rows])
    c_paths = (ctypes.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    _lib.check(_lib.load().sis_png_write_pairs(gen.ctypes.data, lab.ctypes.data, gen.shape[1], gen.shape[2], lab.shape[2], gen.shape[3],
                                               c_rows, c_paths, len(paths), level, max(1, int(n_threads))))
    return paths


# --------------------------------------------------------------------------- running ids across ranks

def assign_round_ids(kept_counts: Sequence[int], n_before: int, num_images: int) -> Tuple[List[Optional[int]], int, bool]:
    """Replay the reference's sequential loop over one round of batches (round t holds batch t*W + r of rank r, in rank
    order).  Returns (first image id of every batch, or None for a batch the reference would never have generated
    because the target was already reached; the new running count; whether the run is finished)."""
    starts: List[Optional[int]] = []
    n = n_before
    for kept in kept_counts:
        if n >= num_images:              # `while pbar.n < args.num_images` (:128)
            starts.append(None)
            continue
        starts.append(n)
        n += int(kept)
    return starts, n, n >= num_images


_HOST_GROUP = None


def exchange_kept_counts(kept: int, device=None) -> List[int]:
    """All-gather of this rank's kept count (one int64 per rank and round); [kept] without a process group.
    The exchange runs over a host-side (gloo) group created on first use: a NCCL collective would be enqueued behind the
    generator passes already queued on the device and stall the loop for a whole batch every round."""
    import torch
    import torch.distributed as dist
    global _HOST_GROUP
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(kept)]
    if dist.get_backend() == 'gloo':
        group = None
    else:
        if _HOST_GROUP is None:
            _HOST_GROUP = dist.new_group(backend='gloo')        # collective: every rank reaches its first `add` together
        group = _HOST_GROUP
    mine = torch.tensor([int(kept)], dtype=torch.int64)
    gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, mine, group=group)
    return [int(g.item()) for g in gathered]


def exchange_kept_counts_async(kept: int):
    """The same all-gather, started now and finished later: returns `wait() -> [kept count of every rank]`.  The writer
    starts the exchange of a round when its batch arrives and only needs the answer when it names that batch's files,
    one batch later: the ranks no longer walk in lock step."""
    import torch
    import torch.distributed as dist
    global _HOST_GROUP
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return lambda: [int(kept)]
    if dist.get_backend() == 'gloo':
        group = None
    else:
        if _HOST_GROUP is None:
            _HOST_GROUP = dist.new_group(backend='gloo')
        group = _HOST_GROUP
    mine = torch.tensor([int(kept)], dtype=torch.int64)
    gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    work = dist.all_gather(gathered, mine, group=group, async_op=True)

    def wait():
        work.wait()
        return [int(g.item()) for g in gathered]
    return wait


class DatasetWriter:
    """The tail of build_dataset's loop for one rank: drop, assign ids, write the PNGs."""

    def __init__(self, base_dir, num_images: int, rank: int = 0, world_size: int = 1, pool=None, device=None,
                 native: Optional[bool] = None, native_threads: Optional[int] = None):
        """`native` (default: whenever PNG_ENCODER is 'fast' or 'stored'): batches go to the library's native writer
        (`sis_png_write_pairs`) from ONE background thread, with `native_threads` encoder threads inside the call (default:
        the size of `pool`, else the host's cores); otherwise one `save_image` task per file on `pool`."""
        self.base_dir, self.num_images = Path(base_dir), num_images
        self.rank, self.world_size, self.pool, self.device = rank, world_size, pool, device
        self.n = 0                       # the reference's pbar.n: images kept so far, over all ranks
        self.finished = False
        self._futures = []
        self.files_written = 0
        self.native = PNG_ENCODER in ('fast', 'stored') if native is None else bool(native)
        self._level = 0 if PNG_ENCODER == 'stored' else 1
        if native_threads is None:
            import os
            native_threads = getattr(pool, '_max_workers', None) or os.cpu_count() or 4
        self.native_threads = native_threads
        self._native_queue = None
        self._awaiting = []              # native path: batches whose kept-count exchange is still in flight
        self.seconds_waiting_for_ranks = 0.0

    def add(self, generated_images: numpy.ndarray, label_images: numpy.ndarray, image_ids_to_drop: Sequence[int]) -> int:
        """One batch of this rank (one round): returns how many files it queued."""
        drop = list(image_ids_to_drop)
        if self.native:
            dropped = set(int(d) for d in drop)
            rows = [b for b in range(len(label_images)) if b not in dropped]
            # The ids of this round need every rank's kept count: the exchange starts now and is waited for when the NEXT
            # batch arrives (single process: at once).  `finished` therefore turns true one round late in multi-rank runs;
            # the extra batch gets no ids (assign_round_ids) and is not written.
            self._awaiting.append((generated_images, label_images, rows, exchange_kept_counts_async(len(rows))))
            queued = 0
            while len(self._awaiting) > (1 if self.world_size > 1 else 0):
                queued += self._name_and_write(self._awaiting.pop(0))
            return queued
        generated_images = numpy.delete(generated_images, drop, axis=0)
        label_images = numpy.delete(label_images, drop, axis=0)
        counts = exchange_kept_counts(len(label_images), self.device)
        starts, self.n, self.finished = assign_round_ids(counts, self.n, self.num_images)
        start = starts[self.rank if len(starts) > 1 else 0]
        if start is None or len(label_images) == 0:
            return 0
        out = save_generated_images(generated_images, label_images, start, self.base_dir, self.num_images, self.pool)
        if self.pool is not None:
            self._futures.extend(out)
        self.files_written += len(out)
        return len(out)

    def _name_and_write(self, item) -> int:
        import time
        generated_images, label_images, rows, wait = item
        t0 = time.perf_counter()
        counts = wait()
        self.seconds_waiting_for_ranks += time.perf_counter() - t0
        starts, self.n, self.finished = assign_round_ids(counts, self.n, self.num_images)
        start = starts[self.rank if len(starts) > 1 else 0]
        if start is None or not rows:
            return 0
        if self._native_queue is None:
            from concurrent.futures import ThreadPoolExecutor
            # two batches in flight: a batch's files rarely divide evenly over the encoder threads, the second call
            # fills the idle tail of the first (the parallelism proper is inside the call)
            self._native_queue = ThreadPoolExecutor(2)
        self._futures.append(self._native_queue.submit(save_generated_images_native, generated_images, label_images, rows, start,
                                                       self.base_dir, self.num_images, self._level, self.native_threads))
        self.files_written += len(rows)
        return len(rows)

    def flush(self):
        while self._awaiting:
            self._name_and_write(self._awaiting.pop(0))
        for f in self._futures:
            f.result()
        self._futures = []
        if self._native_queue is not None:
            self._native_queue.shutdown(wait=True)
            self._native_queue = None


# --------------------------------------------------------------------------- train / val JSON

def iter_through_images_in(image_root: Path, extension: str = 'png') -> Iterable[Path]:
    """coco_gt.py:137-140 (glob order, unsorted, as the reference)."""
    yield from Path(image_root).glob(f'**/*.{extension}')


def determine_classes_in_image(image, class_to_color_map: Dict) -> Dict[str, bool]:
    """coco_gt.py:51-65 on a PIL image or an [S, 2S, 3] array: has_<class> for every non-background class."""
    data = numpy.array(image)
    _, label_image = numpy.split(data, 2, axis=1)
    colors = BaseDatasetSegmenter.load_class_to_color_map(class_to_color_map)
    present = {}
    for class_name, color in colors.items():
        if class_name == 'background':
            continue
        class_mask = numpy.multiply.reduce(label_image[:, :] == color, axis=2)
        contours, _ = cv2.findContours(class_mask.astype('uint8'), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        present[f'has_{class_name}'] = any(c.size >= 6 for c in contours)      # extract_rle keeps contours with >= 3 points
    return present


def create_dataset_json_data(image_paths: Sequence[Path], image_root: Path, class_to_color_map: Dict) -> List[dict]:
    """:151-167."""
    out = []
    for path in image_paths:
        with Image.open(str(path)) as im:
            entry = {'file_name': str(Path(path).relative_to(image_root))}
            entry.update(determine_classes_in_image(im, class_to_color_map))
        out.append(entry)
    return out


def write_train_val_split(image_root, class_to_color_map: Dict, seed: int) -> Tuple[Path, Path]:
    """main, :181-202: shuffle the PNG list under random.seed(seed), first 90 % -> train.json, rest -> val.json."""
    image_root = Path(image_root)
    images = list(iter_through_images_in(image_root))
    random.seed(seed)
    random.shuffle(images)
    split = int(len(images) * 0.9)
    names = []
    for name, part in (('train.json', images[:split]), ('val.json', images[split:])):
        with (image_root / name).open('w') as f:
            json.dump(create_dataset_json_data(part, image_root, class_to_color_map), f)
        names.append(image_root / name)
    return names[0], names[1]


def build_dataset(pair_generator, base_dir, num_images: int, contour_pool=None, writer_pool=None, depth: int = 2,
                  device_contours: Optional[bool] = None) -> Dict:
    """build_dataset's loop (:109-148) on the pipelined B200 path: `pair_generator` is a LabelledPairGenerator; the
    contour stage runs on the device (`device_contours`, see LabelledPairGenerator.iter_segmented; `contour_pool` then
    only serves the rare host fall-backs) or as host tasks on `contour_pool`; PNG writes on `writer_pool`.
    Returns counters."""
    writer = DatasetWriter(base_dir, num_images, pair_generator.rank, pair_generator.world_size, writer_pool,
                           device=pair_generator.generator.input.input.device if pair_generator.world_size > 1 else None)
    import time
    batches, t_add, t0 = 0, 0.0, time.perf_counter()
    for sb in pair_generator.iter_segmented(depth=depth, pool=contour_pool, device_contours=device_contours):
        t1 = time.perf_counter()
        writer.add(sb.images, sb.label_images, sb.image_ids_to_drop)
        t_add += time.perf_counter() - t1
        batches += 1
        if writer.finished:
            break
    t_loop = time.perf_counter() - t0
    writer.flush()
    t_flush = time.perf_counter() - t0 - t_loop
    return {'images_kept_all_ranks': writer.n, 'files_written_this_rank': writer.files_written, 'batches_this_rank': batches,
            'contour_stage': dict(getattr(pair_generator, 'contour_stats', None) or {'host': True}),
            'seconds': {'loop': round(t_loop, 4), 'writer_add': round(t_add, 4), 'final_flush': round(t_flush, 4),
                        'waiting_for_ranks': round(writer.seconds_waiting_for_ranks, 4)}}
