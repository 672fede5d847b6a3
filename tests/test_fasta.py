"""FASTA (last row): the oracle's restatement of entrypos_fasta against the golden vectors made
by the unmodified reference, and the CUDA chain against the oracle."""
import base64
import json
import os
import random

import numpy as np
import pytest

import fqgen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def kat():
    d = json.load(open(os.path.join(GOLD, 'fasta_kat.json')))
    d['blobs'] = [base64.b64decode(b) for b in d['blobs']]
    return d


def test_oracle_entrypos_fasta_matches_reference_vectors(oracle, kat):
    for bi, off, st, *pos in kat['calls']:
        got = [-7] * 4
        assert oracle.entrypos_fasta(kat['blobs'][bi], off, got) == st, (kat['blobs'][bi], off)
        assert got == pos, (kat['blobs'][bi], off)
    # the reference's own FASTA cases (tests.py:83-107): statuses 6 / 3 / 2 for NOTFINAL / FINAL / NOSEQ
    want = {'NOTFINAL': 6, 'FINAL': 3, 'NOSEQ': 2}
    for t in kat['upstream_templates']:
        blob = base64.b64decode(t['blob'])
        pos = [-1] * 4
        st = oracle.entrypos_fasta(blob, 0, pos)
        if t['template'] == 'NOSEQ' or b'\nA' in blob:  # NOTFINAL / FINAL with an empty sequence are not upstream cases
            assert st == want[t['template']], blob
        if st in (3, 6):
            assert blob[pos[0] + 1:pos[1]] == b'foo#2'


def test_oracle_fasta_chain_matches_reference_vectors(oracle, kat):
    for c in kat['chains']:
        blob = kat['blobs'][c['blob']]
        rows, st, tail, resume = oracle.fasta_chain(blob)
        assert rows.tolist() == c['rows'], blob
        assert st == c['status'] and resume == c['offset'], blob
        assert [p if p >= 0 else -7 for p in tail.tolist()] == c['pos'], blob


def test_oracle_against_reference_live(oracle):
    ref = oracle.reference()
    if ref is None:
        pytest.skip('oracle/_ref not present')
    rng = random.Random(99)
    for _ in range(300):
        blob = fqgen.fasta_bytes(rng)
        for off in (0, rng.randrange(0, len(blob) + 1)):
            a, b = [-7] * 6, [-7] * 4
            assert ref[0].entrypos_fasta(blob, off, a) == oracle.entrypos_fasta(blob, off, b)
            assert a[:4] == b
