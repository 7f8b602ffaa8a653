#!/usr/bin/env python
"""bench.py — labelled image pairs/sec at 256^2 (BASELINE.json metric), one JSON line on rank 0.

A step = one pass of the hot path over one batch of synthetic input: StyleGAN2-256 generator forward (random-init
weights, noise/bias parameters perturbed so every path is live) with all 14 activation captures, followed by the
nearest-centroid labelling of layers 8, 9, 12, 13 (k=4, 3 classes) into 256x256 class masks
(BASELINE.json configs[1]: batch 32 on one B200).  N ranks = N independent batch shards (weak scaling); the only
collective is the final all-reduce of the statistics vector.

  value      pairs/s with the step's inputs (latents, noise) already resident in HBM
  e2e        same metric through the public API with HOST buffers: pinned-host latents copied in every step, noise
             drawn on the device as the reference does, image + masks copied back to pinned host memory every step
  roofline   dominant kernel (tcgen05 modulated-conv GEMM): algorithmic conv FLOPs / its CUDA-event time
  cpu_baseline  the oracle (CPU restatement of the reference) on this box's host cores, bounded sample
  --impl reference   the oracle alone, all host threads (the reference has no CPU path of its own)
  gpu_reference      (N=1) the reference's graph on this GPU as the reference runs it, SURVEY §8(d)
  --leg contours | dataset_gan   stage benchmarks of the rows after the hot path, own JSON line
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SIZE, STYLE_DIM, N_MLP, BATCH = 256, 512, 8, 32
LABEL_LAYERS = {'8': 4, '9': 4, '12': 4, '13': 4}
CLASS_MAP = {'0': 'background', '1': 'printed_text', '2': 'handwritten_text', '3': 'background'}
COLORS = {'background': '#000000', 'printed_text': '#0000FF', 'handwritten_text': '#FF0000'}
METRIC = 'labelled image pairs/sec at 256^2'
UNIT = 'pairs/s'


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_burst': d['bf16_tflops'], 'bf16_sustained': d['bf16_tflops_sustained'], 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_burst': 1590.0, 'bf16_sustained': 1400.0, 'source': 'fallback'}


def synthetic_catalog(seed=5):
    """unit-norm centroids, torch.manual_seed(5); F.normalize(randn(k, C))."""
    from oracle import stylegan2_oracle as so
    ch = so.get_channels(2)
    g = torch.Generator().manual_seed(seed)
    cat = {}
    for layer, k in LABEL_LAYERS.items():
        res = 4 if int(layer) <= 1 else 2 ** ((int(layer) - 2) // 2 + 3)
        cat[layer] = torch.nn.functional.normalize(torch.randn(k, ch[res], generator=g), dim=1)
    return cat


def oracle_state():
    from oracle import stylegan2_oracle as so
    spec = so.GeneratorSpec(SIZE, STYLE_DIM, N_MLP, 2)
    sd = so.perturb_zero_params(so.init_state_dict(spec, seed=0), seed=1234)
    return spec, sd


def oracle_step(spec, sd, catalog, inv_map, batch):
    """One CPU pass of the reference's path (generator + labelling) on `batch` samples."""
    from oracle import labelling_oracle as lo
    from oracle import stylegan2_oracle as so
    z = torch.randn(batch, STYLE_DIM)
    noise = so.make_noise(spec)
    img, acts = so.generator_forward(sd, spec, [z], noise=noise, return_intermediate_activations=True)
    masks = lo.prepare_image_segmentation(acts, catalog, inv_map, SIZE)
    return img, masks


def time_oracle(steps, warmup, batch):
    from oracle import labelling_oracle as lo
    torch.set_num_threads(os.cpu_count() or 1)
    spec, sd = oracle_state()
    catalog = synthetic_catalog()
    inv = lo.invert_class_label_map({layer: CLASS_MAP for layer in LABEL_LAYERS})
    torch.manual_seed(1)
    for _ in range(warmup):
        oracle_step(spec, sd, catalog, inv, batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(spec, sd, catalog, inv, batch)
    dt = time.perf_counter() - t0
    return steps * batch / dt, dt / steps * 1e3


def load_ref_kernels():
    """The reference's own two CUDA ops compiled by oracle/build_ref.py (None when oracle/_ref is absent)."""
    import importlib.util
    mods = {}
    for name in ('ref_fused', 'ref_upfirdn2d'):
        path = os.path.join(ROOT, 'oracle', '_ref', f'{name}.so')
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


def time_reference_gpu(dev, steps, warmup, batch):
    """SURVEY.md §8(d) "GPU reference path": the reference's graph on THIS GPU, as the reference runs it -- cuDNN/cuBLAS
    fp32 convs (TF32 off), its own fused_bias_act / upfirdn2d kernels (oracle/_ref), ~150 launches per forward, then
    FactorCatalog.predict with its CPU round trip (factor_catalog.py:47-62: A.cpu(), distances + argmin on the host,
    ids .cuda()), class merge and nearest resize on the GPU.  A reported baseline (bench leg), never the product path."""
    from oracle import labelling_oracle as lo
    from oracle import stylegan2_oracle as so
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_num_threads(os.cpu_count() or 1)
    spec, sd = oracle_state()
    sd = {k: v.to(dev) for k, v in sd.items()}
    catalog = synthetic_catalog()
    inv = lo.invert_class_label_map({layer: CLASS_MAP for layer in LABEL_LAYERS})
    ref = load_ref_kernels()
    saved = (so.fused_bias_act, so.upfirdn2d_op, lo.predict)

    def predict_round_trip(x, centroids):
        b, _, h, w = x.shape
        flat = lo.partial_flat(x).cpu()
        ids = torch.argmin(lo.pairwise_distances(flat, centroids.cpu()), dim=1)
        return ids.to(x.device).reshape(b, h, w)

    def step():
        z = torch.randn(batch, STYLE_DIM).to(dev)
        noise = [n.to(dev) for n in so.make_noise(spec)]
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        with torch.no_grad():
            img, acts = so.generator_forward(sd, spec, [z], noise=noise, return_intermediate_activations=True)
            torch.cuda.synchronize(dev)
            t1 = time.perf_counter()
            masks = lo.prepare_image_segmentation(acts, catalog, inv, SIZE)
        torch.cuda.synchronize(dev)
        return t1 - t0, time.perf_counter() - t0

    try:
        if ref is not None:
            so.fused_bias_act = ref['ref_fused'].fused_bias_act
            so.upfirdn2d_op = ref['ref_upfirdn2d'].upfirdn2d
        lo.predict = predict_round_trip
        torch.manual_seed(1)
        for _ in range(warmup):
            step()
        fwd = tot = 0.0
        for _ in range(steps):
            a, b = step()
            fwd += a
            tot += b
    finally:
        so.fused_bias_act, so.upfirdn2d_op, lo.predict = saved
    return {'value': steps * batch / tot, 'unit': UNIT, 'forward_only_images_per_s': steps * batch / fwd,
            'ms_per_step': tot / steps * 1e3, 'forward_ms_per_step': fwd / steps * 1e3,
            'kind': 'reference graph on this GPU: cuDNN/cuBLAS fp32 (TF32 off) + '
                    + ("the reference's compiled fused_bias_act/upfirdn2d kernels" if ref is not None else 'torch restatement of its two ops')
                    + ', labelling through its CPU round trip (factor_catalog.py:47-62)',
            'cores': torch.get_num_threads(), 'sample': f'{steps} timed + {warmup} warm-up steps of {batch} images, wall clock around synchronize'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [l for (t, l) in self.lines if t_begin <= t <= t_end] or [l for (_, l) in self.lines]
        for l in rows:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


def make_config(world, B):
    return {'workload': 'StyleGAN2 256x256 random-init generator, batch 32 per GPU, 14 activation captures, '
                        'nearest-centroid labelling of layers 8,9,12,13 (k=4) to 256x256 class masks (BASELINE configs[1])',
            'batch_per_gpu': B, 'global_batch': B * world, 'image_size': SIZE, 'k': 4, 'label_layers': list(LABEL_LAYERS),
            'parallelism': f'batch-index sharding over {world} GPU(s), no data-path collective',
            'l2': 'per-step working set (>= 3.9 GB of captured activations) exceeds the 126 MB L2; no explicit flush'}


def run_reference(args, rank, world, out):
    """Reference arm: the oracle port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    batch = 1
    value, ms = time_oracle(args.steps, max(1, min(args.warmup, 2)), batch)
    cores = torch.get_num_threads()
    sample = f'each step = {batch} image(s) (not the 32 of the GPU arm) of the 256^2 config (generator + labelling of layers 8,9,12,13), fp32, {cores} threads'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': make_config(max(1, world), BATCH),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), file=out, flush=True)


def run_leg_contours(args, out):
    """`--leg contours`: CPU timing of the contour stage on synthetic document-like masks: product
    (synthesis_in_style_b200/contours.py) vs the restatement of the reference's algorithm (oracle/contour_oracle.py),
    same inputs, results asserted equal.  No GPU needed."""
    from concurrent.futures import ProcessPoolExecutor

    import numpy

    from oracle import contour_oracle as co
    from synthesis_in_style_b200 import contours as pc
    colors = {'background': (0, 0, 0), 'printed_text': (0, 0, 255), 'handwritten_text': (255, 0, 0)}
    images_n, oracle_n, size = 16, 4, SIZE
    workers = min(8, os.cpu_count() or 1)
    pred = co.synthetic_document_masks(21, images_n, size)
    cfg = pc.ContourConfig(size, colors, ['8', '9'], ['12', '13'], True, 10)
    pc.segment_masks({k: {n: m[:1] for n, m in v.items()} for k, v in pred.items()}, 1, cfg)     # warm-up
    t0 = time.perf_counter()
    images, drop = pc.segment_masks(pred, images_n, cfg)
    t_prod = time.perf_counter() - t0
    with ProcessPoolExecutor(workers) as pool:
        for _ in range(3):
            pc.segment_masks_parallel(pred, images_n, cfg, pool)                                  # workers start lazily
        t0 = time.perf_counter()
        for _ in range(3):
            images_p, drop_p = pc.segment_masks_parallel(pred, images_n, cfg, pool)
        t_par = (time.perf_counter() - t0) / 3
    assert numpy.array_equal(images, images_p) and sorted(drop) == sorted(drop_p)
    sub = {k: {nm: m[:oracle_n] for nm, m in v.items()} for k, v in pred.items()}
    t0 = time.perf_counter()
    o_images, o_drop = co.create_segmentation_image(sub, oracle_n, size, colors, ['8', '9'], ['12', '13'], True, 10)
    t_or = time.perf_counter() - t0
    assert numpy.array_equal(o_images, images[:oracle_n]) and sorted(o_drop) == sorted(d for d in drop if d < oracle_n)
    print(json.dumps({'leg': 'contours', 'image_size': size, 'images': images_n,
                      'product_ms_per_image_1_core': round(t_prod / images_n * 1e3, 2),
                      'product_images_per_s_pool': round(images_n / t_par, 1), 'pool_workers': workers,
                      'reference_algorithm_ms_per_image_1_core': round(t_or / oracle_n * 1e3, 1),
                      'speedup_1_core': round((t_or / oracle_n) / (t_prod / images_n), 1), 'results_equal': True}), file=out, flush=True)


def run_leg_dataset_gan(args, out):
    """`--leg dataset_gan`: throughput of the DatasetGAN labeller at the BASELINE shape: 256^2,
    14 captures (F = 5888), 3 networks, 3 classes; captures from the B200 generator.  Timed with CUDA events:
    sis_pixel_ensemble_label over a batch and its per-category split; beside it the reference's algorithm in PyTorch on
    the same device (materialised [B,S,S,F] features, three fp32 MLPs, TF32 off) on a small batch, labels compared."""
    from oracle import dataset_gan_oracle as dg
    from oracle import stylegan2_oracle as so
    from synthesis_in_style_b200 import _lib
    from synthesis_in_style_b200 import dataset_gan as pg
    from synthesis_in_style_b200.model import Generator
    assert torch.cuda.is_available(), 'this leg needs a CUDA device'
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    batch, ref_batch, iters = args.batch, 2, 5
    spec, sd = oracle_state()
    g = Generator(SIZE, STYLE_DIM, N_MLP)
    g.load_state_dict(sd)
    g = g.to(dev).eval()
    torch.manual_seed(1)
    with torch.no_grad():
        _, acts = g([torch.randn(batch, STYLE_DIM).to(dev)], return_intermediate_activations=True, noise=[n.to(dev) for n in so.make_noise(spec)])
    feat = sum(t.shape[1] for t in acts.values())
    states = [dg.init_classifier_state(feat, 3, seed=50 + i, base_seed=49) for i in range(3)]
    ens = pg.PixelEnsembleClassifier(3, 0, 0)
    for st in states:
        net = pg.PixelClassifier(3, feat)
        net.load_state_dict(st)
        ens.add_network(net.eval())
    for _ in range(2):
        labels, _, _ = ens.predict_label_images(acts, SIZE)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    _lib.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        labels, _, _ = ens.predict_label_images(acts, SIZE)
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    ens.check(dev)
    ms = e0.elapsed_time(e1) / iters
    models = [dg.ClassifierParams({k: v.to(dev) for k, v in st.items()}) for st in states]
    sub = {k: v[:ref_batch] for k, v in acts.items()}
    with torch.no_grad():
        dg.predict_labels(models, sub, SIZE)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        want, margin, _ = dg.predict_labels(models, sub, SIZE)
        r1.record()
        torch.cuda.synchronize()
    ref_ms = r0.elapsed_time(r1)
    safe = margin > 1e-3
    got = labels[:ref_batch].float()
    flops = 2.0 * batch * 384 * sum(t.shape[1] * t.shape[-1] ** 2 for t in acts.values())
    print(json.dumps({'leg': 'dataset_gan', 'image_size': SIZE, 'batch': batch, 'features': feat, 'networks': 3,
                      'ms_per_batch': round(ms, 3), 'images_per_s': round(batch / ms * 1e3, 1),
                      'ms_by_category': {k: round(v[0] / iters, 3) for k, v in prof.items() if v[1]},
                      'first_layer_alg_TFLOP/s': round(flops / (prof['conv_tc'][0] / iters * 1e-3) / 1e12, 1),
                      'reference_algorithm_same_gpu': {'batch': ref_batch, 'ms_per_image': round(ref_ms / ref_batch, 2),
                                                       'images_per_s': round(ref_batch / ref_ms * 1e3, 1),
                                                       'note': 'materialised [B,S,S,F] features + three fp32 MLPs in PyTorch (TF32 off)'},
                      'labels_equal_where_margin_gt_1e-3': bool((got[safe] == want[safe]).all()),
                      'label_agreement': round(float((got == want).float().mean()), 6)}), file=out, flush=True)


def claim_stdout():
    """Native libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout.  Keep a
    private handle on the real stdout for that line and point fd 1 at stderr for everything else."""
    real = os.fdopen(os.dup(1), 'w')
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'fp32'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-steps', type=int, default=3)
    ap.add_argument('--in-flight', type=int, default=2, choices=[1, 2],
                    help='batches in flight on separate CUDA streams / generator workspaces (3 measured slower: do not)')
    ap.add_argument('--leg', default='', choices=['', 'contours', 'dataset_gan'],
                    help='run one of the extra stage benchmarks (rows after the hot path) instead of the headline metric')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world, out)
        return
    if args.leg:
        if rank == 0:
            (run_leg_contours if args.leg == 'contours' else run_leg_dataset_gan)(args, out)
        return

    import torch.distributed as dist
    from synthesis_in_style_b200 import _lib, dataset_creation as dc, labelling
    from synthesis_in_style_b200.model import Generator
    from oracle import stylegan2_oracle as so   # bench's cpu_baseline leg + shared synthetic-weight recipe

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    B = args.batch
    spec, sd = oracle_state()
    g = Generator(SIZE, STYLE_DIM, N_MLP, precision=args.precision)
    g.load_state_dict(sd)
    g = g.to(dev).eval()
    catalog = {k: labelling.FactorCatalog(v.shape[0], v) for k, v in synthetic_catalog().items()}
    seg = labelling.ClusterSegmenter(None, SIZE, COLORS, keys_for_class_determination=['8', '9'],
                                     keys_for_finegrained_segmentation=['12', '13'], num_clusters=4, keys_to_merge={},
                                     catalog=catalog, class_label_map={layer: CLASS_MAP for layer in LABEL_LAYERS})
    cfg = {'batch_size': B, 'latent_size': STYLE_DIM}
    total_steps = args.warmup + args.steps
    # this rank's shard of the reference's single (latent, noise) stream: batch index = rank + i*world
    stream = dc.sharded_latent_stream(g, cfg, seed=1, rank=rank, world_size=world)
    batches = [next(stream)[1].to(dev) for _ in range(total_steps)]

    # `--in-flight` lanes: independent batches alternate over CUDA streams, each lane with its own generator workspace
    # (a replica: same weights, separate native plan), exactly what LabelledPairGenerator(in_flight=...) does
    import copy
    n_lanes = max(1, args.in_flight)
    gens = [g] + [copy.deepcopy(g).eval() for _ in range(n_lanes - 1)]
    lane_streams = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)] if n_lanes > 1 else [None]
    lane_out = [None] * n_lanes

    def step_resident(lat, lane=0):
        # one native call per batch: generator + in-forward labelling (fused with ToRGB where both read the same tensor)
        jobs = seg.make_label_jobs(gens[lane], B)
        with torch.no_grad():
            img, acts = gens[lane]([lat.latent], noise=lat.noise, return_intermediate_activations=True, label_jobs=jobs)
        return img, seg.jobs_to_stacked(jobs)

    def run_steps(first, count):
        cur = torch.cuda.current_stream(dev)
        if n_lanes == 1:
            for i in range(count):
                lane_out[0] = step_resident(batches[first + i])
            return
        for st in lane_streams:
            st.wait_stream(cur)
        done = []
        for i in range(count):
            lane = i % n_lanes
            if i >= n_lanes:
                done[i - n_lanes].synchronize()     # one batch in flight per lane: the lanes stay half a step apart
            with torch.cuda.stream(lane_streams[lane]):
                lane_out[lane] = step_resident(batches[first + i], lane)
                ev = torch.cuda.Event()
                ev.record(lane_streams[lane])
                done.append(ev)
        for st in lane_streams:
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- value: inputs resident in HBM
    run_steps(0, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    run_steps(args.warmup, args.steps)
    stats = dc.reduce_stats(torch.cat([seg.cluster_pixel_counts[k] for k in sorted(seg.cluster_pixel_counts)]).clone())
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = _lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = world * args.steps * B / (ms_total / 1e3)

    # ---------------------------------------------------------------- e2e: host buffers in and out
    # Public API: LabelledPairGenerator.iter_host — pinned-host latents in, device noise as the reference draws it,
    # fp32 image + 4 x [3, B, S, S] uint8 mask stacks out to pinned host memory every step (copies on a side stream).
    n_cls = len(COLORS)
    h2d = B * STYLE_DIM * 4
    d2h = B * 3 * SIZE * SIZE * 4 + len(LABEL_LAYERS) * n_cls * B * SIZE * SIZE
    pipe = dc.LabelledPairGenerator(g, seg, cfg, seed=1, rank=rank, world_size=world, in_flight=n_lanes)
    host_iter = pipe.iter_host(depth=2)
    for _ in range(8):                               # past the one-time costs (replica plans, pinned slots): steady state
        next(host_iter)
    barrier()
    checksum = 0
    e0.record()
    for _ in range(args.steps):
        hb = next(host_iter)
        checksum += int(hb.masks['13'][1, 0, 0, 0])      # touch the host result of every step
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * B / (float(ms2.item()) / 1e3)

    # ---------------------------------------------------------------- roofline: per-kernel CUDA-event times
    roofline, kernels = None, None
    if rank == 0:
        peaks = load_peaks()
        _lib.profile_enable(True)
        torch.cuda.synchronize()
        for i in range(args.profile_steps):
            step_resident(batches[args.warmup + (i % args.steps)])
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        n = args.profile_steps
        flops_img = so.conv_flops_per_image(spec)
        rgb_flops = 2.0 * sum((2 ** i) ** 2 * spec.channels[2 ** i] * 3 for i in range(2, spec.log_size + 1))
        conv_flops_step = (flops_img - rgb_flops) * B
        kernels = {k: {'ms_per_step': v[0] / n, 'launches_per_step': v[1] / n} for k, v in prof.items() if v[1]}
        conv_key = 'conv_tc' if args.precision == 'bf16x3' else 'conv_simt'
        conv_ms = kernels[conv_key]['ms_per_step']
        achieved = conv_flops_step / (conv_ms / 1e3) / 1e12
        peak = peaks['bf16_sustained']
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(conv_key)
        roofline = {'bound': 'tensor', 'kernel': 'modconv_tc_kernel' if conv_key == 'conv_tc' else 'modconv3x3_simt_kernel',
                    'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                    'peak_source': f'{peaks["source"]} bf16 sustained (kernel timed inside a long step)',
                    'passes': 3 if conv_key == 'conv_tc' else 1,
                    'note': 'algorithmic FLOPs (2*MACs, SURVEY 8d) of the 13 StyledConv layers (14 GEMM launches) per step / their summed '
                            'CUDA-event time; the bf16x3 split issues 3 MMA passes per algorithmic FLOP, so tensor-pipe '
                            'work is 3x achieved',
                    'share_of_step': conv_ms / sum(v['ms_per_step'] for v in kernels.values())}
        # HBM-side view of the memory-bound kernels (bytes model in DESIGN.md)
        hbm = {}
        act_bytes = sum(B * spec.channels[4 if i <= 1 else 2 ** ((i - 2) // 2 + 3)] * (4 if i <= 1 else 2 ** ((i - 2) // 2 + 3)) ** 2 * 4
                        for i in range(spec.n_latent))
        # labelling and ToRGB: the 64^2 and 256^2 ToRGBs ride in the labelling passes of layers 9 and 13 (one read of those tensors)
        lbl_bytes = sum(B * spec.channels[r] * r * r * 4 + n_cls * B * SIZE * SIZE for r in (64, 64, 256, 256))
        rgb_bytes = sum(B * spec.channels[r] * r * r * 4 + B * 3 * r * r * 4 + B * 3 * (r // 2) ** 2 * 4 for r in (4, 8, 16, 32, 64, 128, 256))
        fused_saved = B * spec.channels[256] * 256 * 256 * 4 + B * spec.channels[64] * 64 * 64 * 4
        lr_ms = kernels.get('label', {'ms_per_step': 0})['ms_per_step'] + kernels.get('torgb', {'ms_per_step': 0})['ms_per_step']
        if lr_ms > 0:
            hbm['label+torgb'] = {'GB/s': (lbl_bytes + rgb_bytes - fused_saved) / (lr_ms / 1e3) / 1e9,
                                  'note': '4 labelling launches (two of them with the ToRGB of the same tensor fused in) + 5 ToRGB launches; '
                                          'the tiny 4^2..32^2 ToRGB launches are latency-bound'}
        if 'blur_split' in kernels:
            bl_bytes = sum(B * spec.channels[r] * ((r + 1) ** 2 * 4 + r * r * 8) for r in (8, 16, 32, 64, 128, 256))
            hbm['blur_split'] = {'GB/s': bl_bytes / (kernels['blur_split']['ms_per_step'] / 1e3) / 1e9}
        for v in hbm.values():
            v['frac_of_hbm_peak'] = v['GB/s'] / peaks['hbm_gbs']
        roofline['memory_bound_kernels'] = hbm
        roofline['captured_activation_bytes_per_step'] = act_bytes

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms_cpu = time_oracle(3, 1, 1)
        cpu = {'value': v, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '3 timed + 1 warm-up steps of 1 image (256^2 generator + labelling of layers 8,9,12,13), oracle fp32'}

    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpu_ref = time_reference_gpu(dev, 2, 1, B)

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16x3 (3-term bf16 split, fp32 accumulate)' if args.precision == 'bf16x3' else 'f32', 'data': 'synthetic',
                'config': dict(make_config(world, B), in_flight_batches=n_lanes,
                               pipelining=f'{n_lanes} independent batches in flight per GPU on separate CUDA streams / generator workspaces'),
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': float(ms2.item()) / args.steps},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'kernels': kernels, 'cpu_baseline': cpu,
                'gpu_reference': gpu_ref, 'stats_allreduce_sum': int(stats.sum().item())}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
