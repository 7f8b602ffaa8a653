#!/usr/bin/env python
"""Summarise ncu outputs (read here, no GPU) into small tracked files under profiles/.
  launches CSV (gpu__time_duration.sum per launch)  -> per-kernel totals / shares of the profiled command
  .ncu-rep full captures                             -> one row per launch with the metrics the roofline uses
Usage: scripts/summarize_ncu.py <tag>   (reads gpurun_out/*_<tag>.*, writes profiles/<tag>_*.md)
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name.replace('sis::', '')[:90]


def launches():
    path = os.path.join(G, f'launches_{tag}.csv')
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(io.StringIO(''.join(lines))))
    per = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        v_us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
        k = short(r['Kernel Name'])
        d = per.setdefault(k, [0, 0.0])
        d[0] += 1; d[1] += v_us
        total += v_us
    out = [f'# ncu launch list `{tag}` — `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1`',
           '', '`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes).',
           f'{sum(d[0] for d in per.values())} launches, {total / 1e3:.2f} ms total.', '',
           '| kernel | launches | total us | share |', '|---|---:|---:|---:|']
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append(f'| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f} % |')
    open(os.path.join(P, f'{tag}_launches.md'), 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:30]))


METRICS = OrderedDict([
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_%'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1_%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_%'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
])


def reps():
    for f in sorted(os.listdir(G)):
        if f.endswith(f'_{tag}.ncu-rep'):
            if os.path.exists(os.path.join(G, f.replace('.ncu-rep', '.csv'))):
                continue
            text = subprocess.run(['ncu', '-i', os.path.join(G, f), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        elif f.startswith('prof_') and f.endswith(f'_{tag}.csv') and '_source_' not in f:
            text = open(os.path.join(G, f)).read()
        else:
            continue
        rows = list(csv.reader(io.StringIO(text)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        out = [f'# ncu --set full capture `{f}`', '', '`--clock-control none --import-source on`; one row per launch; '
               'units as reported by ncu (' + ', '.join(f'{v}: {units[col[k]]}' for k, v in METRICS.items() if k in col and units[col[k]]) + ').', '',
               '| kernel | ' + ' | '.join(METRICS.values()) + ' |', '|---|' + '---:|' * len(METRICS)]
        for r in rows[2:]:
            vals = [r[col[k]] if k in col else '-' for k in METRICS]
            vals = [f'{float(v):.3f}' if re.match(r'^-?\d+\.\d+$', v) else v for v in vals]
            out.append(f'| `{short(r[col["Kernel Name"]])}` | ' + ' | '.join(vals) + ' |')
        if f.startswith('prof_conv_') and 'last' not in f and len(rows) - 2 in (13, 14):
            # DRAM traffic of the conv GEMM launches of one step -> bench.py's roofline.traffic
            import json
            def gb(r, k):
                v, u = float(r[col[k]]), units[col[k]]
                return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}.get(u, 1.0)
            tot = sum(gb(r, 'dram__bytes_read.sum') + gb(r, 'dram__bytes_write.sum') for r in rows[2:])
            json.dump({'conv_tc': tot, 'unit': 'bytes per step (all conv GEMM launches of one step, B=32, 256^2)', 'source': f'profiles/{f.replace(".csv", ".md")}'},
                      open(os.path.join(P, 'traffic.json'), 'w'))
        name = f.replace('.ncu-rep', '.md').replace('.csv', '.md')
        open(os.path.join(P, name), 'w').write('\n'.join(out) + '\n')
        print('\n'.join(out))


def source_pages(top=30):
    """`ncu --page source --csv` of a capture taken with --import-source on: line 1 names the kernel, line 2 is the
    header, then one row per SASS instruction.  Keeps the `top` instructions by warp-stall samples with their opcode text,
    share of all samples, executed count and dominant stall reason."""
    for f in sorted(os.listdir(G)):
        if not (f.startswith('prof_') and '_source_' in f and f.endswith(f'_{tag}.csv')):
            continue
        rows = list(csv.reader(open(os.path.join(G, f))))
        if len(rows) < 3:
            continue
        kernel = rows[0][1] if len(rows[0]) > 1 else '?'
        hdr = rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        s_col = col.get('# Samples', col.get('Warp Stall Sampling (All Samples)'))
        stall_cols = [(h, i) for h, i in col.items() if h.startswith('stall_') and '(Not Issued)' not in h]

        def num(r, i):
            try:
                return float(r[i].replace(',', '')) if i is not None and i < len(r) and r[i] not in ('', '-') else 0.0
            except ValueError:
                return 0.0
        body = [r for r in rows[2:] if len(r) > 2]
        total = sum(num(r, s_col) for r in body) or 1.0
        ranked = sorted(body, key=lambda r: -num(r, s_col))[:top]
        out = [f'# ncu source page `{f}`', '', f'kernel: `{short(kernel)}`; {len(body)} SASS instructions, {int(total)} warp-stall samples; '
               f'top {len(ranked)} instructions by samples (share of all samples, executions, dominant stall reason).', '',
               '| # | address | SASS | samples | share | executed | top stall |', '|---:|---|---|---:|---:|---:|---|']
        for n, r in enumerate(ranked, 1):
            stalls = sorted(((num(r, i), h) for h, i in stall_cols), reverse=True)
            top_stall = f'{stalls[0][1]} ({stalls[0][0]:.0f})' if stalls and stalls[0][0] > 0 else '-'
            sass = re.sub(r'\s+', ' ', r[col['Source']]).strip()[:90]
            out.append(f'| {n} | {r[col["Address"]][-6:]} | `{sass}` | {int(num(r, s_col))} | {100 * num(r, s_col) / total:.1f} % | '
                       f'{int(num(r, col.get("Instructions Executed")))} | {top_stall} |')
        by_op = {}
        for r in body:
            op = r[col['Source']].strip().split(' ')[0].lstrip('@!P0123456789 ') or r[col['Source']].strip().split(' ')[0]
            toks = [t for t in r[col['Source']].strip().split(' ') if t and not t.startswith('@')]
            op = toks[0].split('.')[0] if toks else '?'
            by_op[op] = by_op.get(op, 0.0) + num(r, s_col)
        out += ['', 'Samples by opcode: ' + ', '.join(f'`{k}` {100 * v / total:.1f} %' for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:12])]
        open(os.path.join(P, f.replace('.csv', '.md')), 'w').write('\n'.join(out) + '\n')
        print('\n'.join(out[:12]))


launches()
reps()
source_pages()
