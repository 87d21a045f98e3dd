#!/usr/bin/env python
"""Per-instruction warp-stall samples of an ncu report (SASS view): python tools/ncu_stalls.py rep.ncu-rep [kernel-substring] [top N]
Totals by stall reason, then the instructions that collect the most samples with their dominant reasons."""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(io.StringIO(out)))
i = 0
seen = set()
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i + 1]; i += 2
        body = []
        while i < len(rows) and rows[i] and rows[i][0] != 'Kernel Name':
            body.append(rows[i]); i += 1
        if want not in name or name in seen:
            continue
        seen.add(name)
        print('==', name[:120])
        isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
        reasons = [(k, hdr.index(k)) for k in hdr if k.startswith('stall_') and '(Not Issued)' not in k]
        tot = {k: 0 for k, _ in reasons}
        tsamp = 0; tinst = 0
        for r in body:
            tsamp += int(r[isamp] or 0); tinst += int(r[iex] or 0)
            for k, j in reasons:
                tot[k] += int(r[j] or 0)
        print('  samples', tsamp, 'warp instructions', tinst)
        print('  by reason:', ', '.join(f'{k[6:]} {100 * v / max(tsamp, 1):.1f}%' for k, v in sorted(tot.items(), key=lambda x: -x[1]) if v > 0.005 * tsamp))
        order = sorted(range(len(body)), key=lambda n: -int(body[n][isamp] or 0))[:top]
        for n in sorted(order):
            r = body[n]
            rs = sorted(((int(r[j] or 0), k[6:]) for k, j in reasons), reverse=True)[:3]
            print(f'  {n:4d} {100 * int(r[isamp] or 0) / max(tsamp, 1):5.1f}%  {r[1].strip()[:70]:70s} ' + ' '.join(f'{k}:{v}' for v, k in rs if v))
    else:
        i += 1
