#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Value')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= mi: continue
    name = r[ki].split('(')[0].replace('void ', '')
    tot[name] += float(r[mi].replace(',', '')); cnt[name] += 1
s = sum(tot.values())
print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k, v in tot.most_common():
    print(f"{k:58s} {cnt[k]:8d} {v / 1e6:10.3f} {v / cnt[k] / 1e3:9.1f} {100 * v / s:6.1f}%")
