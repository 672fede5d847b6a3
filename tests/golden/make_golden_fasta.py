"""Golden vectors for FASTA: the UNMODIFIED reference's entrypos_fasta (src/fastqandfurious.py:103-143).

Run in the build container (needs /root/reference through oracle/_ref):

    make -C oracle ref && python tests/golden/make_golden_fasta.py

Output (committed): fasta_kat.json
    calls   single calls: blob, offset -> status, posbuffer (initialised to -7: untouched entries stay -7)
    chains  whole blobs walked with the reference function itself, each call starting at pos3 of the
            previous record: rows, status / posbuffer / offset of the first call that was not COMPLETE
The templates of the reference's own FASTA tests (tests.py:37-57, 83-107) are part of `calls`.
"""
import base64
import json
import os
import random
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import fqgen  # noqa: E402
import oracle  # noqa: E402

oracle.build()
mod, _ = oracle.reference()


def b64(b):
    return base64.b64encode(bytes(b)).decode()


def call(blob, offset):
    pos = [-7] * 6
    st = mod.entrypos_fasta(blob, offset, pos)
    return st, pos[:4]


def chain(blob):
    rows, offset = [], 0
    while True:
        st, pos = call(blob, offset)
        if st != mod.COMPLETE:
            return rows, st, pos, offset
        rows.append(pos)
        offset = pos[3]


def main():
    rng = random.Random(20261017)
    blobs = []
    header, seq, mseq = 'foo#2', 'AATTGCCG', 'AATTGCCG\nGCCGTA'
    tpl = {'NOTFINAL': "\n>{header}\n{sequence}\n>{header}_2\n{sequence}\n", 'FINAL': "\n>{header}\n{sequence}\n",
           'NOSEQ': "\n>{header}\n"}
    upstream = []
    for name, t in tpl.items():
        for s in (seq, mseq, ''):
            blob = t.format(header=header, sequence=s).encode('ascii')
            upstream.append({'template': name, 'blob': b64(blob)})
            blobs.append(blob)
    blobs += [b'', b'\n', b'>', b'\n>', b'\n>\n', b'\n>a', b'\n>a\n', b'\n>a\nA', b'\n>a\nA\n', b'\n>a\n\n>b\nC\n',
              b'\n>a\n>b\nACGT\n>c\nG\n', b'\n>\n>\n>\n>\n>\n', b'\n>>\n>>\n', b'x\n>a\nAC\nGT\n>b\nA', b'\n>a\r\nAC\r\n>b\r\nGT\r\n']
    for _ in range(200):
        blobs.append(fqgen.fasta_bytes(rng))
    calls, chains = [], []
    for bi, blob in enumerate(blobs):
        offs = {0, 1, len(blob) // 2, max(0, len(blob) - 2), len(blob), len(blob) + 3}
        offs |= {rng.randrange(0, len(blob) + 1) for _ in range(3)}
        for off in sorted(offs):
            st, pos = call(blob, off)
            calls.append([bi, off, st] + pos)  # blob index, offset, status, pos0..pos3
        rows, st, pos, off = chain(blob)
        chains.append({'blob': bi, 'rows': rows, 'status': st, 'pos': pos, 'offset': off})
    out = {'reference': 'lgautier/fastq-and-furious src/fastqandfurious.py:103-143 (entrypos_fasta), run unmodified',
           'blobs': [b64(b) for b in blobs], 'upstream_templates': upstream, 'calls': calls, 'chains': chains}
    with open(os.path.join(HERE, 'fasta_kat.json'), 'w') as fh:
        json.dump(out, fh)
    print('calls', len(calls), 'chains', len(chains), 'records', sum(len(c['rows']) for c in chains))


if __name__ == '__main__':
    main()
