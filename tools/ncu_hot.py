#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (one kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = his[which]
end = his[which + 1] - 1 if which + 1 < len(his) else len(rows)
hdr = rows[hi]
print(rows[hi - 1][:2])
data = [r for r in rows[hi + 1:end] if len(r) >= len(hdr) - 2]
isamp, isrc, iex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print(f"{len(data)} instructions, {tot} samples")
top = sorted(range(len(data)), key=lambda k: -int(data[k][isamp] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for k in sorted(top):
    r = data[k]
    s = int(r[isamp] or 0)
    why = sorted(((int(r[i] or 0), h[6:]) for i, h in stalls), reverse=True)[:2]
    print(f"{k:6d} {s:7d} {100.0 * s / tot:5.1f}%  exec {r[iex]:>9s}  {r[isrc].strip()[:70]:70s} {why}")
if len(sys.argv) > 4:
    # sample totals for index ranges a:b,c:d,...
    for rg in sys.argv[4].split(','):
        a, b = [int(x) for x in rg.split(':')]
        tot_r = sum(int(data[k][isamp] or 0) for k in range(a, min(b, len(data))))
        ex = sum(int(data[k][iex] or 0) for k in range(a, min(b, len(data))))
        agg = {}
        for k in range(a, min(b, len(data))):
            for i, h in stalls:
                agg[h[6:]] = agg.get(h[6:], 0) + int(data[k][i] or 0)
        top3 = sorted(agg.items(), key=lambda kv: -kv[1])[:5]
        print(f"range {a}:{b}: {tot_r} samples ({100.0 * tot_r / tot:.1f}%), {ex} warp-instructions; {top3}")
