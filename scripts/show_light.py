#!/usr/bin/env python
"""Print one line per launch of a light ncu CSV (scripts/gpu_profile_light.sh).  Usage: scripts/show_light.py <csv> [name filter]"""
import csv, io, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.reader(io.StringIO(''.join(lines))))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
flt = sys.argv[2] if len(sys.argv) > 2 else ''
def val(r, m):
    v = r[col[m]].replace(',', '')
    return float(v) if v not in ('', '-') else 0.0
for r in rows[2:]:
    name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('sis::', '').replace('void ', '')
    if flt and flt not in name:
        continue
    u = units[col['gpu__time_duration.sum']]; t = val(r, 'gpu__time_duration.sum')
    t_ms = t / 1e6 if u.startswith('n') else (t / 1e3 if u.startswith('u') else t)
    def byt(m):
        return val(r, m) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Tbyte': 1e12}.get(units[col[m]], 1)
    print(f"  {name[:50]:50s} {t_ms:7.3f} ms tensor {val(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):5.1f}% dram {val(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}% l2 {val(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}% l1 {val(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}% rd {byt('dram__bytes_read.sum')/1e9:.3f} wr {byt('dram__bytes_write.sum')/1e9:.3f} GB grid {val(r,'launch__grid_size'):.0f}")
