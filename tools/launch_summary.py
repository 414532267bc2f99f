#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr, data = rows[hi], rows[hi + 1:]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= iv:
        continue
    name = r[ik].split('(')[0].replace('hcu::', '').replace('void ', '')
    v = float(r[iv].replace(',', ''))
    v = v / 1000 if r[iu] == 'ns' else (v * 1000 if r[iu] == 'ms' else v)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"{sum(cnt.values())} launches, {s:.1f} us total (cold-cache, serialised: compare shares)")
for n, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{v:10.1f} us {cnt[n]:4d}x  {v / s * 100:5.1f}%  {n[:90]}")
