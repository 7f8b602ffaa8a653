#!/usr/bin/env python
"""Print the figures of a bench.py JSON line (headline + extra configs) in a few lines.  Usage: scripts/show_bench.py <file>"""
import json
import sys

d = json.loads(open(sys.argv[1]).read())


def one(tag, c):
    if 'error' in c:
        print(tag, 'ERROR', c['error'])
        return
    r = c.get('roofline') or {}
    k = c.get('kernels') or {}
    par = c.get('parity') or {}
    print(f"{tag}: value {c['value']:.0f} pairs/s  {c['ms_per_step']:.3f} ms/step  e2e {c['e2e']['value']:.0f}  launches {c['gpu_launches']}  "
          f"parity ok={par.get('ok')} agree={par.get('label_agreement', 0):.6f} bad={par.get('label_mismatches_at_margin_gt_1e-3')} err={par.get('image_max_abs_err', 0):.1e}")
    print('   kernels(ms): ' + ', '.join(f"{n} {v['ms_per_step']:.3f}" for n, v in k.items()))
    if r:
        extra = ''
        if 'narrow_layers' in r:
            extra = f"  wide {r['wide_layers']['TFLOP/s']:.0f} TF/s  narrow {r['narrow_layers']['TFLOP/s']:.0f} TF/s ({r['narrow_layers']['achieved']:.0f} GB/s = {r['narrow_layers']['frac']:.2f} of HBM)"
        mb = r.get('memory_bound_kernels', {})
        print(f"   conv {r['achieved']:.0f} TF/s alg, frac {r['frac']:.3f}, of split cap {r.get('frac_of_split_cap', 0):.3f}{extra}; "
              + ', '.join(f"{n} {v['frac_of_hbm_peak']:.2f} of HBM" for n, v in mb.items()))


one('config ' + str(d['config'].get('baseline_config')), d)
print('   clocks', d.get('clocks'))
for c in d.get('configs', []):
    one(f"config {c.get('config')}", c)
if d.get('cpu_baseline'):
    print('   cpu_baseline', round(d['cpu_baseline']['value'], 2), 'gpu_ref', round(d['gpu_reference']['value'], 2), 'tf32', round(d['gpu_reference_tf32']['value'], 2))
