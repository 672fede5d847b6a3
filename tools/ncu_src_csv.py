#!/usr/bin/env python
"""Summarise a SASS source page exported on the GPU box (ncu -i X.ncu-rep --page source --csv > X.csv):
instructions executed and stall samples per SASS instruction.  usage: tools/ncu_src_csv.py X.csv [top N | all]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hdr = next(r for r in rows if r and r[0] == 'Address')
k = [r for r in rows if len(r) == len(hdr) and r is not hdr]
iS, iI, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iI] or 0) for r in k)
tots = sum(int(r[iSamp] or 0) for r in k)
print('total warp instructions', tot, 'samples', tots, 'n sass', len(k))
agg = {h: sum(int(r[i] or 0) for r in k) for i, h in stalls}
print('stall samples:', ', '.join('%s %.1f%%' % (h[6:], 100.0 * v / max(1, tots)) for h, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 100 > tots))
mode = sys.argv[2] if len(sys.argv) > 2 else 'all'
for n, r in enumerate(k):
    c = int(r[iI] or 0)
    sm = int(r[iSamp] or 0)
    if mode != 'all' and sm * 200 < tots and c * 200 < tot:
        continue
    top = max(stalls, key=lambda ih: int(r[ih[0]] or 0))
    print('%4d %10d %5.2f%% %6d %5.2f%% %-14s %s' % (n, c, 100.0 * c / tot, sm, 100.0 * sm / max(1, tots), top[1][6:] if sm else '', r[iS].strip()[:100]))
