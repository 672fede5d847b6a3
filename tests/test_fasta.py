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


@pytest.fixture(scope='module')
def fq():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g
    g.build()
    import fastqandfurious_b200 as m
    return m


def _dev(data, offset=0):
    import torch
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    t = torch.empty(len(a) + offset + 32, dtype=torch.uint8, device='cuda')
    t.fill_(10 if offset else ord('>'))  # hostile neighbours
    if len(a):
        t[offset:offset + len(a)].copy_(torch.from_numpy(a.copy()))
    return t[offset:offset + len(a)]


def _check_chain(fq, oracle, data, sentinel, goff, offset=0, **kw):
    blob = (b'\n' if sentinel else b'') + bytes(data)
    want, st, tail, resume = oracle.fasta_chain(blob, 0, goff)
    res = fq.parse_fasta_buffer(_dev(data, offset), sentinel=bool(sentinel), goff=goff, **kw)
    ctx = (bytes(data)[:120], sentinel, goff, offset)
    assert res.n == len(want), ctx
    assert np.array_equal(res.table.cpu().numpy(), want), ctx
    assert res.tail_status == st and res.tail_pos == tail.tolist() and res.resume_offset == resume, ctx
    return res


@pytest.mark.gpu
def test_fasta_chain_golden_blobs(fq, oracle, kat):
    """Every golden blob through the CUDA chain: rows, status, positions and offset of the first call that is
    not COMPLETE equal the unmodified reference's (the chains in fasta_kat.json were walked by it)."""
    for c in kat['chains']:
        blob = kat['blobs'][c['blob']]
        res = fq.parse_fasta_buffer(_dev(blob), sentinel=False, goff=0)
        assert res.table.cpu().numpy().tolist() == c['rows'], blob
        assert res.tail_status == c['status'] and res.resume_offset == c['offset'], blob
        assert [p if p >= 0 else -7 for p in res.tail_pos] == c['pos'], blob


@pytest.mark.gpu
@pytest.mark.parametrize('seed', range(3))
def test_fasta_chain_random(fq, oracle, seed):
    rng = random.Random(500 + seed)
    for _ in range(120):
        data = fqgen.fasta_bytes(rng)
        for sentinel in (1, 0):
            # (both scan geometries of the FASTA parse: 32 KiB per iteration, the default, and 16 KiB)
            _check_chain(fq, oracle, data, sentinel, rng.choice([-1, 0, 1000]), offset=rng.randrange(16), cfg=rng.choice([0, 1]))


@pytest.mark.gpu
def test_fasta_large_and_runs(fq, oracle):
    rng = random.Random(77)
    # many tiles, wrapped sequences, records across tile borders
    big = b''.join(b'>r%d desc\n' % k + fqgen._wrap(bytes(rng.choice(b'ACGT') for _ in range(rng.randint(0, 900))), 60) + b'\n'
                   for k in range(4000))
    res = _check_chain(fq, oracle, big, 1, -1)
    assert res.n == 3999 and res.tail_status == 3
    _check_chain(fq, oracle, big[:-1], 1, -1)
    _check_chain(fq, oracle, big, 1, -1, cfg=1)
    _check_chain(fq, oracle, big, 1, -1, max_lines=16)   # workspace too small at first: retried with the reported need
    # long runs of header-only records (every other one is swallowed as "sequence"), also across tile borders
    runs = b'>h\n' * 30000 + b'>x\nACGT\n' + b'>\n' * 7 + b'>y\nAC\n>z\n'
    _check_chain(fq, oracle, runs, 1, -1)
    _check_chain(fq, oracle, runs, 0, 0)
    # runs that cover whole tiles (run lengths come from per-tile summaries, not from a walk): dense and sparse
    for head in (b'', b'>a\nAC\n', b'AC\n'):
        _check_chain(fq, oracle, head + b'>\n' * 400_001 + b'>y\nAC\n>z\n', 1, -1)
    sparse = b''.join(b'>' + b'h' * rng.randint(1, 50_000) + b'\n' for _ in range(301)) + b'>y\nAC\n>z\n'
    _check_chain(fq, oracle, sparse, 1, -1)
    _check_chain(fq, oracle, b'ACGT\n' + sparse, 0, 0)
    # one unwrapped sequence spanning hundreds of tiles between two records
    long_seq = b'>chr1\n' + bytes(rng.choice(b'ACGT') for _ in range(3_000_000)) + b'\n>chr2\nACGT\n'
    res = _check_chain(fq, oracle, long_seq, 1, -1)
    assert res.n == 1 and int(res.table[0, 3] - res.table[0, 2]) == 3_000_000


@pytest.mark.gpu
def test_entrypos_fasta_plugin_and_upstream_cases(fq, oracle, kat):
    """The per-call adaptor against the reference's single-call vectors, and the reference's own FASTA test
    (tests.py:83-107) run against it with the entryfunc_fasta it expects."""
    from array import array
    for bi, off, st, *pos in kat['calls'][::7]:
        got = array('q', [-7] * 6)
        blob = kat['blobs'][bi]
        assert fq.entrypos_fasta(blob, off, got) == st, (blob, off)
        assert list(got[:4]) == pos, (blob, off)
    header, seq, mseq = 'foo#2', 'AATTGCCG', 'AATTGCCG\nGCCGTA'
    cases = (("\n>{header}\n{sequence}\n>{header}_2\n{sequence}\n", fq.COMPLETE, seq),
             ("\n>{header}\n{sequence}\n>{header}_2\n{sequence}\n", fq.COMPLETE, mseq),
             ("\n>{header}\n{sequence}\n", fq.MISSING_SEQ_END, seq), ("\n>{header}\n{sequence}\n", fq.MISSING_SEQ_END, mseq),
             ("\n>{header}\n", fq.MISSING_SEQ_BEG, ''))
    for tpl, status, s in cases:
        entries = tpl.format(header=header, sequence=s).encode('ascii')
        posbuffer = array('q', [-1] * 6)
        assert fq.entrypos_fasta(entries, 0, posbuffer) == status
        h, q = fq.entryfunc_fasta(entries, posbuffer, 0)
        assert h == header.encode('ascii')
        if status != fq.MISSING_SEQ_BEG:
            assert q == s.encode('ascii')
