#!/usr/bin/env python
"""Per-source-line totals from an ncu report: python tools/ncu_lines.py rep.ncu-rep [kernel-substring]
(instructions executed per warp-launch, stall samples, shared/global wavefront columns where present)."""
import csv, subprocess, sys, io
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ''
rows = list(csv.reader(io.StringIO(out)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Function Name':
        name = rows[i][1]; hdr = rows[i + 1]; i += 2
        show = want in name
        if show: print('==', name[:110])
        ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
        tot_e = tot_s = 0
        lines = []
        while i < len(rows) and rows[i] and rows[i][0] not in ('Function Name', 'File Path'):
            r = rows[i]
            if r[0] != '' and len(r) > ie:
                lines.append((int(r[0]), r[1].strip(), int(r[ie] or 0), int(r[isamp] or 0)))
            i += 1
        if show:
            te = sum(l[2] for l in lines) or 1; ts = sum(l[3] for l in lines) or 1
            for ln, src, e, s in lines:
                if e > 0.004 * te or s > 0.004 * ts:
                    print(f'{ln:5d} {100*e/te:5.1f}%inst {100*s/ts:5.1f}%stall  {src[:100]}')
            print('  total inst', te, 'samples', ts)
    else:
        i += 1
