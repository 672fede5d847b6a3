"""Long differential run of the oracle's C restatement against the compiled, unmodified reference (oracle/_ref) on
the CPU: single entrypos / entrypos_fasta calls at random offsets and whole streams through readfastq_iter with random
fbufsize.  Needs /root/reference to have been compiled (oracle.build()).  usage: python tests/fuzz_oracle.py [seconds]"""
import io
import os
import random
import sys
import time
from array import array

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))  # fqgen, algo_model

import numpy as np  # noqa: E402
import fqgen  # noqa: E402
import oracle  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
oracle.build()
ref = oracle.reference()
assert ref is not None, 'oracle/_ref not built'
mod, cext = ref
rng = random.Random(int(os.environ.get('FUZZ_SEED', '5')))
t0 = time.time()
n_calls = n_streams = n_fasta = 0
seed = 0
while time.time() - t0 < seconds:
    seed += 1
    for d in fqgen.corpus(100000 + seed, 50):
        if rng.random() < 0.3:
            d = fqgen.mutate(rng, d, rng.randint(1, 4))
        blob = b'\n' + d
        for off in {0, rng.randrange(len(blob) + 1), rng.randrange(len(blob) + 1)}:
            if blob[off:].endswith(b'\n@') or blob.endswith(b'\n@'):
                continue  # out-of-bounds read in the reference (memchr length -1): not reproduced
            want, got = array('q', [0] * 6), array('q', [0] * 6)
            a, b = cext.entrypos(blob, off, want), oracle.entrypos(blob, off, got)
            assert a == b and list(want) == list(got), (blob, off, a, b, list(want), list(got))
            n_calls += 1
        # whole stream: rows and error of readfastq_iter + C entrypos + entryfunc_abspos, any fbufsize
        if b'\n@' not in (d[-2:], d[-1:] + b'@') and not d.endswith(b'\n@'):
            fb = rng.choice([1, 7, 100, 65536])
            table, oerr, obyte = oracle.readfastq(d)
            if oerr == 3:
                continue  # the reference loops forever on INVALID (src/fastqandfurious.py:256-270); the oracle reports it
            rows, err = [], None
            try:
                it = mod.readfastq_iter(io.BytesIO(d), fb, entryfunc=mod.entryfunc_abspos, entrypos=cext.entrypos)
                for k, p in enumerate(it):
                    rows.append(list(p))
                    if k > 5000:
                        break
            except ValueError as e:
                err = str(e)
            except Exception:  # the reference's own loop can die on damaged input; not comparable
                continue
            want_err = {0: None, 1: 'Incomplete final quality string at byte', 2: 'Incomplete entry at byte %i' % obyte}[oerr]
            assert table.tolist() == rows and err == want_err, (d, fb, err, want_err, len(rows), len(table))
            n_streams += 1
    for _ in range(50):
        blob = fqgen.fasta_bytes(rng)
        for off in (0, rng.randrange(0, len(blob) + 1)):
            a, b = [-7] * 6, [-7] * 4
            assert mod.entrypos_fasta(blob, off, a) == oracle.entrypos_fasta(blob, off, b) and a[:4] == b, (blob, off)
            n_fasta += 1
print('oracle vs compiled reference: %d entrypos calls, %d streams, %d entrypos_fasta calls in %.0f s: all identical'
      % (n_calls, n_streams, n_fasta, time.time() - t0))
