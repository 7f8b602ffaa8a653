#!/usr/bin/env python
"""Generator forward timing at the other BASELINE shapes (configs[2] 512^2 batch 16, configs[3] 1024^2 batch 8), with the
library's per-category CUDA-event split.  One JSON line per shape."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesis_in_style_b200 import _lib  # noqa: E402
from synthesis_in_style_b200.model import Generator  # noqa: E402


def conv_flops(g, size):
    ch = g.channels
    total, res, cin = 16 * 9 * ch[4] * ch[4], 4, ch[4]
    while res < size:
        res *= 2
        total += (res // 2) ** 2 * 9 * cin * ch[res] + res * res * 9 * ch[res] * ch[res]
        cin = ch[res]
    return 2.0 * total


def main():
    dev = torch.device('cuda:0')
    for size, batch in ((256, 32), (512, 16), (1024, 8)):
        torch.manual_seed(0)
        g = Generator(size, 512, 8).to(dev).eval()
        z = torch.randn(batch, 512, device=dev)
        noise = g.make_noise()
        with torch.no_grad():
            for _ in range(3):
                g([z], noise=noise, return_intermediate_activations=True)
            torch.cuda.synchronize()
            _lib.profile_enable(True); _lib.profile_collect()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 5
            e0.record()
            for _ in range(iters):
                g([z], noise=noise, return_intermediate_activations=True)
            e1.record(); torch.cuda.synchronize()
            prof = _lib.profile_collect(); _lib.profile_enable(False)
        ms = e0.elapsed_time(e1) / iters
        conv_ms = prof['conv_tc'][0] / iters
        fl = conv_flops(g, size) * batch
        print(json.dumps({'size': size, 'batch': batch, 'ms_per_forward': round(ms, 3), 'images_per_s': round(batch / ms * 1e3, 1),
                          'ms_by_category': {k: round(v[0] / iters, 3) for k, v in prof.items() if v[1]},
                          'conv_alg_TFLOP/s': round(fl / (conv_ms * 1e-3) / 1e12, 1)}), flush=True)
        del g
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
