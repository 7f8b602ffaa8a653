#!/usr/bin/env python
"""BASELINE.json configs[4], modulated_conv2d leg: the stand-alone ModulatedConv2d (plain & up-sampling, demodulation on /
off) over C in {64..512} and res in {8..512}, timed per category with the library's CUDA-event profiler, so the
GEMM kernel time is separated from the operand pre-scale / blur passes around it.
Prints one JSON line per case: algorithmic TFLOP/s (2*MACs, reference's transposed-conv count for `up`) of the GEMM
launch, its fraction of the measured bf16 peak (x3 MMA passes stated), and the whole-op time.
Usage: scripts/conv_sweep.py [--only CIN,COUT,RES,UP[,BATCH] ...]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesis_in_style_b200 import _lib  # noqa: E402
from synthesis_in_style_b200.model import ModulatedConv2d  # noqa: E402


def peak_bf16():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['bf16_tflops']
    return 1590.0


NCU = os.environ.get('SIS_SWEEP_NCU') == '1'      # under ncu: one call per case, so launch order = case order


def run_case(dev, cin, cout, res, up, demod, batch, iters=8):
    torch.manual_seed(0)
    warm = 3
    if NCU:
        iters, warm = 1, 0
    m = ModulatedConv2d(cin, cout, 3, 512, demodulate=demod, upsample=up).to(dev)
    nbuf = 3
    xs = [torch.randn(batch, cin, res, res, device=dev) for _ in range(nbuf)]
    style = torch.randn(batch, 512, device=dev)
    with torch.no_grad():
        for i in range(warm):
            m(xs[i % nbuf], style)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        _lib.profile_collect()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            m(xs[i % nbuf], style)
        e1.record()
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
    ms_op = e0.elapsed_time(e1) / iters
    # all GEMM launches of one call (the Cout = 128 up-conv issues two; layers with Cout <= 64 are timed under their own category)
    ms_gemm = (prof['conv_tc'][0] + prof.get('conv_tc_narrow', (0.0, 0))[0]) / iters
    flops = 2.0 * batch * res * res * 9 * cin * cout
    tf = flops / (ms_gemm * 1e-3) / 1e12
    return {'op': 'modulated_conv2d', 'cin': cin, 'cout': cout, 'res_in': res, 'up': up, 'demod': demod, 'B': batch,
            'gemm_ms': round(ms_gemm, 4), 'op_ms_profiled': round(ms_op, 4), 'alg_TFLOP/s': round(tf, 1),
            'mma_pass_TFLOP/s': round(3 * tf, 1), 'frac_bf16_peak_x3': round(3 * tf / peak_bf16(), 3),
            'other_ms': {k: round(v[0] / iters, 4) for k, v in prof.items() if v[1] and k not in ('conv_tc', 'conv_tc_narrow')}}


def main():
    dev = torch.device('cuda:0')
    cases = []
    if '--only' in sys.argv:
        for spec in sys.argv[sys.argv.index('--only') + 1:]:
            f = [int(v) for v in spec.split(',')]
            cin, cout, res, up = f[:4]
            cases.append((cin, cout, res, bool(up), True) + ((f[4],) if len(f) > 4 else ()))
    else:
        for c in (64, 128, 256, 512):
            for res in (8, 32, 64, 128, 256, 512):
                if c * res * res * 4 * 2 > (1 << 30):       # keep one sample's in+out under 1 GiB
                    continue
                cases.append((c, c, res, False, True))
                if res <= 256:
                    cases.append((c, c, res, True, True))
        cases.append((128, 128, 256, False, False))
        cases.append((512, 512, 64, False, False))
        cases.append((512, 256, 64, True, True))
        cases.append((256, 128, 128, True, True))
    for case in cases:
        cin, cout, res, up, demod = case[:5]
        out_res = res * 2 if up else res
        per = (cin * res * res + cout * out_res * out_res * (3 if up else 1)) * 4 + cout * out_res * out_res * 4
        batch = case[5] if len(case) > 5 else int(max(1, min(64, (1 << 30) // per)))
        print(json.dumps(run_case(dev, cin, cout, res, up, demod, batch)), flush=True)


if __name__ == '__main__':
    main()
