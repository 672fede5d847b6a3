#!/usr/bin/env python
"""Summarise an ncu report's SASS source page: instructions executed per SASS instruction, grouped
by loop region; prints the top instructions and totals.  usage: tools/ncu_src.py rep [kernel-idx]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'] , capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
kernels = []
cur = None
for r in rows:
    if r and r[0] == 'Address':
        hdr = r; cur = []; kernels.append(cur); continue
    if hdr and cur is not None and len(r) == len(hdr):
        cur.append(r)
k = kernels[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
iS, iI, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iI] or 0) for r in k)
tots = sum(int(r[iSamp] or 0) for r in k)
print('total warp instructions', tot, 'samples', tots, 'n sass', len(k))
for n, r in enumerate(k):
    c = int(r[iI] or 0)
    print('%4d %10d %5.2f%% %6s  %s' % (n, c, 100.0 * c / tot, r[iSamp], r[iS][:110]))
