#!/usr/bin/env python
"""BASELINE config 5 — kernel sweep of the two native ops (and the labelling kernel) over channels 64-512 and
resolutions 8-1024: achieved HBM GB/s (algorithmic bytes / CUDA-event time) against the measured peak.
Each case uses tensors larger than the 126 MB L2 where the shape allows it (batch chosen to reach ~1 GiB) or
rotates over enough distinct buffers to defeat the cache.  One JSON line per case on stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from synthesis_in_style_b200.op import fused_leaky_relu, upfirdn2d  # noqa: E402
from synthesis_in_style_b200 import labelling  # noqa: E402


def peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    return json.load(open(p))['hbm_gbs'] if os.path.exists(p) else 6650.0


NCU = os.environ.get('SIS_SWEEP_NCU') == '1'      # under ncu: one launch per case, so launch order = case order


def time_op(fn, inputs, iters=10, warm=3):
    """fn(*inputs[i % len(inputs)]); distinct buffers rotate so that every iteration misses L2."""
    if NCU:
        iters, warm = 1, 0
    for i in range(warm):
        fn(*inputs[i % len(inputs)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(*inputs[i % len(inputs)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    dev = torch.device('cuda:0')
    hbm = peak()
    target_bytes = 1 << 30
    k4 = torch.tensor([1., 3., 3., 1.], device=dev)
    k4 = (k4[None] * k4[:, None]) / 64 * 4
    out = []
    for c in (64, 128, 256, 512):
        for res in (8, 32, 64, 128, 256, 512, 1024):
            n_elem = c * res * res
            b = max(1, min(256, target_bytes // (n_elem * 4)))
            if b * n_elem * 4 > (2 << 30):
                continue
            nbuf = max(1, min(8, (400 << 20) // (b * n_elem * 4) + 1))
            # ---- fused_leaky_relu: 2*N*4 + C*4 bytes
            xs = [(torch.randn(b, c, res, res, device=dev), torch.randn(c, device=dev)) for _ in range(nbuf)]
            t = time_op(lambda x, bias: fused_leaky_relu(x, bias), xs)
            by = 2 * b * n_elem * 4 + c * 4
            out.append({'op': 'fused_leaky_relu', 'C': c, 'res': res, 'B': b, 'us': t * 1e6, 'GB/s': by / t / 1e9, 'frac': by / t / 1e9 / hbm})
            # ---- upfirdn2d mode 1 (Blur after the up-conv: (res+1)^2 -> res^2, pad (1,1)): (N_in + N_out)*4 bytes
            if res >= 8:
                xs1 = [(torch.randn(b, c, res + 1, res + 1, device=dev),) for _ in range(nbuf)]
                t = time_op(lambda x: upfirdn2d(x, k4, pad=(1, 1)), xs1)
                by = (b * c * (res + 1) ** 2 + b * n_elem) * 4
                out.append({'op': 'upfirdn2d_blur_mode1', 'C': c, 'res': res, 'B': b, 'us': t * 1e6, 'GB/s': by / t / 1e9, 'frac': by / t / 1e9 / hbm})
                del xs1
            del xs
            torch.cuda.empty_cache()
    # ---- upfirdn2d mode 3 (ToRGB skip upsample: [B,3,res/2,res/2] -> [B,3,res,res], pad (2,1))
    for res in (64, 256, 1024):
        b = max(1, min(512, target_bytes // (3 * res * res * 5)))
        nbuf = 4
        xs = [(torch.randn(b, 3, res // 2, res // 2, device=dev),) for _ in range(nbuf)]
        t = time_op(lambda x: upfirdn2d(x, k4, up=2, pad=(2, 1)), xs)
        by = (b * 3 * (res // 2) ** 2 + b * 3 * res * res) * 4
        out.append({'op': 'upfirdn2d_up_mode3', 'C': 3, 'res': res, 'B': b, 'us': t * 1e6, 'GB/s': by / t / 1e9, 'frac': by / t / 1e9 / hbm})
        del xs
    # ---- labelling (k = 4 and 20) on the default config's shapes
    for (c, res, b) in ((512, 64, 32), (128, 256, 32), (64, 512, 16)):
        for k in (4, 20):
            xs = [(torch.randn(b, c, res, res, device=dev),) for _ in range(3)]
            cent = torch.nn.functional.normalize(torch.randn(k, c, device=dev), dim=1)
            bits = torch.tensor([1 << (i % 3) for i in range(k)], dtype=torch.int32, device=dev)
            size = max(res, 256)
            t = time_op(lambda x: labelling.label_assign(x, cent, class_bits=bits, n_class=3, image_size=size), xs)
            by = b * c * res * res * 4 + 3 * b * size * size + k * c * 4
            out.append({'op': f'label_assign_k{k}', 'C': c, 'res': res, 'B': b, 'us': t * 1e6, 'GB/s': by / t / 1e9, 'frac': by / t / 1e9 / hbm})
            del xs
    for r in out:
        print(json.dumps(r))


if __name__ == '__main__':
    main()
