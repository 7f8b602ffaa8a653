#!/usr/bin/env python
"""Join the ncu pass of the kernel sweeps (one launch of the op per case, in case order) with the event-timed sweep:
profiles/<tag>_kernel_sweep.md -- per case: event-timed GB/s or TFLOP/s, and ncu's time, DRAM throughput % and tensor-pipe %.
Usage: scripts/summarize_sweep_ncu.py <tag>"""
import csv
import io
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def ncu_rows(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.reader(io.StringIO(''.join(lines))))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        def val(m):
            v = r[col[m]].replace(',', '')
            return float(v) if v not in ('', '-') else 0.0
        u = units[col['gpu__time_duration.sum']]
        t = val('gpu__time_duration.sum')
        us = t / 1e3 if u.startswith('n') else (t if u.startswith('u') else t * 1e3)
        out.append({'name': re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('sis::', '').replace('void ', ''), 'us': us,
                    'dram': val('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                    'tensor': val('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), 'grid': int(val('launch__grid_size'))})
    return out


def jsonl(path):
    return [json.loads(l) for l in open(path) if l.strip().startswith('{')]


out = [f'# Kernel sweep `{tag}` (BASELINE config 5)', '',
       'Event columns: CUDA-event timing over 8-10 launches on rotating buffers (`scripts/kernel_sweep.py`, `scripts/conv_sweep.py`). '
       'ncu columns: one launch of the same case under `ncu --clock-control none` (cold, serialised): time, '
       '`gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed`, `sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active`.', '']
mem_cases, mem_ncu = jsonl(os.path.join(G, f'{tag}_kernel_sweep.jsonl')), ncu_rows(os.path.join(G, f'{tag}_kernel_sweep_ncu.csv'))
out += ['## Memory-bound ops', '', '| op | C | res | B | event us | GB/s | of HBM peak | ncu kernel | ncu us | ncu dram % |', '|---|---:|---:|---:|---:|---:|---:|---|---:|---:|']
for i, c in enumerate(mem_cases):
    n = mem_ncu[i] if i < len(mem_ncu) and len(mem_ncu) == len(mem_cases) else None
    out.append(f"| {c['op']} | {c['C']} | {c['res']} | {c['B']} | {c['us']:.1f} | {c['GB/s']:.0f} | {c['frac']:.2f} | "
               + (f"`{n['name'][:40]}` | {n['us']:.1f} | {n['dram']:.1f} |" if n else '- | - | - |'))
if len(mem_ncu) != len(mem_cases):
    out += ['', f'(ncu captured {len(mem_ncu)} launches for {len(mem_cases)} cases: not joined.)']
conv_cases, conv_ncu = jsonl(os.path.join(G, f'{tag}_conv_sweep.jsonl')), ncu_rows(os.path.join(G, f'{tag}_conv_sweep_ncu.csv'))
out += ['', '## modulated_conv2d', '', '| cin | cout | res_in | up | B | GEMM ms (event) | alg TFLOP/s | MMA-pass TFLOP/s | ncu GEMM kernels: us @ tensor % | ncu blur: us @ dram % |',
        '|---:|---:|---:|---|---:|---:|---:|---:|---|---|']
it = iter(conv_ncu)
pending = next(it, None)
for c in conv_cases:
    gemm, blur = [], []
    # one call = its GEMM launch(es), then (up-sampling layers) one blur launch
    while pending is not None and pending['name'].startswith('modconv_tc'):
        gemm.append(pending)
        pending = next(it, None)
        if not c['up'] and gemm:
            break
        if c['up'] and pending is not None and not pending['name'].startswith('modconv_tc'):
            break
    if c['up'] and pending is not None and pending['name'].startswith('blur_act_split'):
        blur.append(pending)
        pending = next(it, None)
    out.append(f"| {c['cin']} | {c['cout']} | {c['res_in']} | {'up' if c['up'] else 'plain'} | {c['B']} | {c['gemm_ms']:.3f} | {c['alg_TFLOP/s']:.0f} | "
               f"{c['mma_pass_TFLOP/s']:.0f} | " + '; '.join(f"`{g['name'][22:48]}` {g['us']:.0f} @ {g['tensor']:.0f}" for g in gemm) + ' | '
               + '; '.join(f"{b['us']:.0f} @ {b['dram']:.0f}" for b in blur) + ' |')
open(os.path.join(P, f'{tag}_kernel_sweep.md'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:60]))
