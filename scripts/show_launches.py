#!/usr/bin/env python
"""Aggregate an ncu launch list (--csv --page raw) by kernel name.  Usage: scripts/show_launches.py <csv> [name filter]"""
import collections, csv, io, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.reader(io.StringIO(''.join(lines))))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
flt = sys.argv[2] if len(sys.argv) > 2 else ''
u = units[col['gpu__time_duration.sum']]
k = 1e-3 if u.startswith('n') else (1.0 if u.startswith('u') else 1e3)
agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('sis::', '').replace('void ', '')
    if flt and flt not in name:
        continue
    t = float(r[col['gpu__time_duration.sum']].replace(',', '')) * k
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] = max(a[2], t)
for name, a in agg.items():
    print(f'{name[:60]:60s} n={a[0]:4d} total={a[1]:10.1f} us  max={a[2]:9.1f} us')
print(f'total {sum(a[1] for a in agg.values()):.1f} us in {sum(a[0] for a in agg.values())} launches')
