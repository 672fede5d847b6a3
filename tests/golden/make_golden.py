"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    make -C oracle ref && python tests/golden/make_golden.py

The reference's compiled C extension (`_fastqandfurious.entrypos`, src/_fastqandfurious.c:25-153)
and its `readfastq_iter` / `entryfunc_abspos` (src/fastqandfurious.py:186-279) are loaded from
oracle/_ref (built by oracle/Makefile from the sources where they lie under /root/reference).
Outputs (committed):
    entrypos_kat.json     single entrypos calls: blob, offset -> status, posbuffer   (C ext and Py)
    readfastq_kat.json    whole streams -> abspos table or the ValueError message, several fbufsize
    arrayadd_kat.json     arrayadd_b / arrayadd_q probes
The three *.fq files are byte copies of the reference's data/ fixtures.
"""
import base64
import io
import json
import os
import random
import sys
from array import array

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import fqgen  # noqa: E402
import oracle  # noqa: E402

oracle.build()
mod, cext = oracle.reference()


def b64(b):
    return base64.b64encode(bytes(b)).decode()


def c_entrypos(blob, offset):
    pos = array('q', [-7] * 6)
    st = cext.entrypos(blob, offset, pos)
    return st, list(pos)


def py_entrypos(blob, offset):
    pos = array('q', [-7] * 6)
    st = mod.entrypos(blob, offset, pos)
    return st, list(pos)


class Loop(Exception):
    pass


def run_stream(data, fbufsize):
    """readfastq_iter + C entrypos + entryfunc_abspos; guards the reference's infinite loop on an
    INVALID entry at end of stream (src/fastqandfurious.py:256-270 falls through every branch)."""
    calls = [0]

    def guarded(buf, offset, posbuffer):
        calls[0] += 1
        if calls[0] > 4 * len(data) + 64:
            raise Loop()
        return cext.entrypos(buf, offset, posbuffer)

    rows = []
    try:
        for pos in mod.readfastq_iter(io.BytesIO(data), fbufsize, entryfunc=mod.entryfunc_abspos,
                                      entrypos=guarded):
            rows.append(list(pos))
    except ValueError as e:
        return {'rows': rows, 'error': str(e)}
    except Loop:
        return {'rows': rows, 'error': 'LOOP'}
    return {'rows': rows, 'error': None}


def ub(blob):
    """Inputs on which the C extension reads out of bounds (memchr with length -1 when the blob ends
    with "\\n@", src/_fastqandfurious.c:70-71): excluded, never reproduced."""
    return blob.endswith(b'\n@')


def main():
    rng = random.Random(20261017)
    # ---- entrypos known answers -------------------------------------------------------------
    blobs = [
        b'\n@foo#2\nAATTGCCG\n+\n3425@!#!\n', b'\n@foo#2\nAATTGCCG\n+\n3425@!#!\n@barfoo#2\n',
        b'\n@foo#2\nAATTGCCG\n+\n', b'', b'\n', b'\n@foo\n', b'\n@foo\nA', b'\n@foo\nACGT\n+',
        b'\n@foo\nACGT\n+\nI', b'\n@foo\nACGT\n+fo\nIIII\n@x\n', b'\n@foo\nACGT\n+foo\nIIII\n@x\n',
        b'\n@foo\nACGT\n+bar\nIIII\n@x\n', b'\n@h\n\n+\n\n@x\n', b'\n@foo#2\nAATT\nGCCG\n+\n3425\n@!#!\n@b\n',
        b'\n@a\r\nAC\r\n+\r\nII\r\n@b\n',
    ]
    # the truncation cases of the reference's tests.py:141-166
    for tmpl in (b'\n@foo#2\nAATTGCCG\n+\n3425@!#!\n', b'\n@foo#2\nAATT\nGCCG\n+\n3425\n@!#!\n'):
        for cut in range(len(tmpl) + 1):
            blobs.append(tmpl[:cut])
    for d in fqgen.corpus(777, 100):
        blobs.append(b'\n' + d)
        blobs.append(d)
    kat = []
    for blob in blobs:
        if ub(blob):
            continue
        offs = {0, 1, len(blob) // 2, max(0, len(blob) - 3)}
        offs |= {rng.randrange(len(blob) + 1) for _ in range(3)}
        for off in sorted(offs):
            if ub(blob[off:]) or ub(blob):
                continue
            st, pos = c_entrypos(blob, off)
            pst, ppos = py_entrypos(blob, off)
            kat.append({'blob': b64(blob), 'offset': off, 'c': [st, pos], 'py': [pst, ppos]})
    json.dump(kat, open(os.path.join(HERE, 'entrypos_kat.json'), 'w'), separators=(',', ':'))

    # ---- stream known answers ---------------------------------------------------------------
    streams = []
    for name in ('test.fq', 'test_longqualityheader.fq', 'test_multiline.fq'):
        streams.append((name, open(os.path.join(HERE, name), 'rb').read()))
    base = b'@r1\nACGT\n+\nIIII\n@r2\nGG\n+\nII'
    for i, tail in enumerate((b'', b'\n', b'\n\n', b'\n\n\n')):
        streams.append(('eof%d' % i, base + tail))
    streams += [
        ('garbage', b'garbage\n' + base + b'\n'), ('empty', b''), ('nl', b'\n'),
        ('trunc_head', b'@r1\nACGT\n+\nIIII\n@r2'), ('trunc_seq', b'@r1\nACGT\n+\nIIII\n@r2\nGG'),
        ('trunc_plus', b'@r1\nACGT\n+\nIIII\n@r2\nGG\n+'),
        ('atqual', b'@r1\nACGT\n+\n@III\n@r2\nAC\n+\n@+\n@r3\nACGTACGT\n+r3\n++++@@@@\n'),
        ('shortqual', b'@r1\nACGT\n+\nII\n@r2\nAC\n+\nII\n@r3\nA\n+\nI\n'),
        ('longqual', b'@r1\nAC\n+\nIIII\n@r2\nAC\n+\nII\n@r3\nA\n+\nI\n'),
        ('badplus', b'@r1\nACGT\n+zz\nIIII\n@r2\nAC\n+\nII\n'),
        ('crlf', b'@a\r\nAC\r\n+\r\nII\r\n@b\r\nAC\r\n+\r\nII\r\n'),
        ('emptyseq', b'@h\n\n+\n\n@x\nA\n+\nI\n'),
    ]
    k = 0
    for d in fqgen.corpus(888, 250):
        streams.append(('corpus%d' % k, d))
        k += 1
    out = []
    for name, data in streams:
        if ub(b'\n' + data):
            continue
        res = {}
        for fb in (1, 7, 100, 200, 600, 700, 65536):
            if fb < 7 and len(data) > 400:
                continue
            # a chunk boundary can expose the UB blob ("...\n@" at the end of a refill): skip those
            res[str(fb)] = run_stream(data, fb) if not any(
                ub(b'\n' + data[:c]) for c in range(fb, len(data) + 1, fb)) else None
        out.append({'name': name, 'data': b64(data), 'res': res})
    json.dump(out, open(os.path.join(HERE, 'readfastq_kat.json'), 'w'), separators=(',', ':'))

    # ---- arrayadd ---------------------------------------------------------------------------
    aa = []
    for vals, v in (([33, 73, 126], -33), ([-128, 127, 0], -33), ([0, 1, 2], 1000), ([5, -5], 127),
                    ([-128, 127], -128), ([], 3)):
        a = array('b', vals)
        if len(a):
            cext.arrayadd_b(a, v)
        aa.append({'kind': 'b', 'in': vals, 'value': v, 'out': list(a)})
    for vals, v in (([0, 29, 30, 115, 118, 203], -1), ([2 ** 62, -2 ** 62, 0], 12345678901),
                    ([-1] * 6, 2 ** 40)):
        a = array('q', vals)
        cext.arrayadd_q(a, v)
        aa.append({'kind': 'q', 'in': vals, 'value': v, 'out': list(a)})
    json.dump(aa, open(os.path.join(HERE, 'arrayadd_kat.json'), 'w'), separators=(',', ':'))
    print('entrypos cases', len(kat), 'streams', len(out), 'arrayadd', len(aa))


if __name__ == '__main__':
    main()
