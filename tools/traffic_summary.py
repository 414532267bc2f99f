#!/usr/bin/env python
"""conv_tc DRAM traffic per launch from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` -> JSON."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
iid, im, iv, iu = hdr.index('ID'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
per = {}
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    u = r[iu].lower()
    if 'byte' in u:
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u]
    per.setdefault(r[iid], {})[r[im]] = v
n = len(per)
rd = sum(p.get('dram__bytes_read.sum', 0) for p in per.values())
wr = sum(p.get('dram__bytes_write.sum', 0) for p in per.values())
out = {"kernel": "hcu_conv_tc_fwd", "launches": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
       "dram_bytes_per_launch": (rd + wr) / max(1, n),
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the conv_tc launches of one steady-state bench step"}
print(json.dumps(out, indent=1))
