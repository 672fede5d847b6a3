"""Consumers of the offset table (index replay, length filter, quality sums): host-side pieces on CPU, the
CUDA kernels against the oracle's restatement of the reference's slicing recipes on the GPU."""
import io
import os
import random
from array import array

import numpy as np
import pytest

import fqgen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GOLD_ROWS = {'test.fq': [[0, 29, 30, 115, 118, 203], [204, 233, 234, 646, 649, 1061], [1062, 1091, 1092, 1225, 1228, 1361],
                         [1362, 1391, 1392, 1446, 1449, 1503]]}


def test_oracle_consumers_match_reference_slices(oracle):
    """The oracle's field recipes against plain Python slicing as the reference's entryfunc does it, and
    (where oracle/_ref is loadable) against the unmodified reference's entryfunc + arrayadd_b."""
    data = open(os.path.join(GOLD, 'test.fq'), 'rb').read()
    table, err, _ = oracle.readfastq(data)
    assert err == 0 and table.tolist() == GOLD_ROWS['test.fq']
    ref = oracle.reference()
    for field in (0, 1, 2):
        out, off = oracle.gather_fields(data, table, field)
        for k, row in enumerate(table):
            want = (data[row[0] + 1:row[1]], data[row[2]:row[3]], data[row[4]:row[5]])[field]
            assert bytes(out[off[k]:off[k + 1]]) == want
            if ref is not None:
                assert want == ref[0].entryfunc(data, array('q', row.tolist()), 0)[field]
    sums = oracle.field_sums(data, table, 2, add=-33)
    for k, row in enumerate(table):
        q = array('b')
        q.frombytes(data[row[4]:row[5]])
        if ref is not None:
            ref[1].arrayadd_b(q, -33)
        else:
            q = array('b', [x - 33 for x in q])
        assert sums[k] == sum(q)
    assert oracle.select_by_length(table, 1, 100, 200).tolist() == [2]
    assert oracle.field_lengths(table, 1).tolist() == [85, 412, 133, 54]


def test_index_file_round_trip():
    """write_index / read_index against the reference's per-record array('q').tofile / fromfile loop
    (src/demo/benchmark.py:57-63, 276-280); host I/O only."""
    from fastqandfurious_b200 import consume
    rows = np.array(GOLD_ROWS['test.fq'], dtype=np.int64)
    fh = io.BytesIO()
    assert consume.write_index(rows, fh) == 4
    ref = io.BytesIO()
    for r in rows:
        ref.write(array('q', r.tolist()).tobytes())
    assert fh.getvalue() == ref.getvalue()
    fh.seek(0)
    assert np.array_equal(consume.read_index(fh), rows)
    with pytest.raises(EOFError):
        consume.read_index(io.BytesIO(fh.getvalue()[:-8]))
    assert consume.read_index(io.BytesIO(b'')).shape == (0, 6)


@pytest.fixture(scope='module')
def fq():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g
    g.build()
    import fastqandfurious_b200 as m
    return m


def _parsed(fq, data):
    import torch
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    res = fq.parse_buffer(d)
    return d, res.table.clone()


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['illumina', 'ont', 'multiline'])
def test_gather_lengths_sums_match_oracle(fq, oracle, kind):
    import torch
    data = fqgen.variable_records_np(400 if kind != 'ont' else 40, 11, kind).tobytes()
    d, table = _parsed(fq, data)
    want_table, err, _ = oracle.readfastq(data)
    host = table.cpu().numpy()
    assert err == 0 and np.array_equal(host, want_table[:len(host)])
    rng = random.Random(7)
    sels = [None, [], [0], list(range(len(host) - 1, -1, -1)), sorted(rng.sample(range(len(host)), len(host) // 3)),
            [rng.randrange(len(host)) for _ in range(50)]]
    for sel in sels:
        st = None if sel is None else torch.tensor(sel, dtype=torch.int64, device='cuda')
        for field, name in enumerate(('header', 'sequence', 'quality')):
            assert np.array_equal(fq.field_lengths(table, name, st).cpu().numpy(), oracle.field_lengths(host, field, sel))
            for add in (0, -33):
                out, off = fq.gather_fields(d, table, name, st, add=add)
                wout, woff = oracle.gather_fields(data, host, field, sel, add=add)
                assert np.array_equal(off.cpu().numpy(), woff), (kind, name, add)
                assert np.array_equal(out.cpu().numpy(), wout), (kind, name, add)
        assert np.array_equal(fq.field_sums(d, table, 'quality', st).cpu().numpy(), oracle.field_sums(data, host, 2, sel))
    # the fused decode of parse_buffer and a decoding gather agree (both are the arrayadd_b recipe)
    res = fq.parse_buffer(d, decode_quality=True)
    q, off = fq.gather_fields(d, table, 'quality', add=-33)
    mirror = res.qual.cpu().numpy().view(np.uint8)
    want = np.concatenate([mirror[r[4]:r[5]] for r in host]) if len(host) else np.empty(0, np.uint8)
    assert np.array_equal(q.cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['illumina', 'ont', 'multiline', 'mixed'])
def test_pack_2bit_matches_oracle(fq, oracle, kind):
    """fqb_pack_2bit against the oracle's definition of the layout: short records (a lane each), long ones (the warp
    shares one), wrapped ones (newlines squeezed out; a long wrapped record is redone by one lane), other letters
    counted, row subsets, a table whose positions are offset from the buffer, an unaligned buffer."""
    import torch
    rng = random.Random(13)
    if kind == 'mixed':
        recs = []
        for k in range(300):
            n = rng.choice([0, 1, 15, 16, 17, 63, 64, 65, 150, 2047, 2048, 2049, 5000, 40000])
            seq = bytes(rng.choice(b'ACGTacgtNnURYKM.*-') if rng.random() < 0.05 else rng.choice(b'ACGT') for _ in range(n))
            wrap = rng.choice([0, 0, 60, 13]) if n else 0
            body = b'\n'.join(seq[i:i + wrap] for i in range(0, n, wrap)) if wrap else seq
            qual = bytes(rng.choice(b'IJK') for _ in range(n))
            qbody = b'\n'.join(qual[i:i + wrap] for i in range(0, n, wrap)) if wrap else qual
            if n == 0:
                continue  # the reference's C entrypos does not parse an empty sequence line
            recs.append(b'@r%d\n' % k + body + b'\n+\n' + qbody + b'\n')
        data = b''.join(recs)
    else:
        data = fqgen.variable_records_np(400 if kind != 'ont' else 40, 12, kind).tobytes()
    want_table, err, _ = oracle.readfastq(data)
    assert err == 0 and len(want_table) > 30
    for shift in (0, 3):
        raw = torch.full((len(data) + shift + 8,), 65, dtype=torch.uint8, device='cuda')
        raw[shift:shift + len(data)].copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        d = raw[shift:shift + len(data)]
        table = fq.parse_buffer(d).table
        host = table.cpu().numpy()
        assert np.array_equal(host, want_table[:len(host)])
        sels = [None, [], [len(host) - 1], sorted(rng.sample(range(len(host)), len(host) // 2)),
                [rng.randrange(len(host)) for _ in range(40)]]
        for sel in sels:
            st = None if sel is None else torch.tensor(sel, dtype=torch.int64, device='cuda')
            packed, off, nb, no = fq.pack_2bit(d, table, st)
            wp, woff, wnb, wno = oracle.pack_2bit(data, host, sel)
            assert np.array_equal(off.cpu().numpy(), woff), (kind, shift)
            assert np.array_equal(nb.cpu().numpy(), wnb) and np.array_equal(no.cpu().numpy(), wno), (kind, shift)
            assert np.array_equal(packed.cpu().numpy(), wp), (kind, shift)
        # table positions relative to a stream of which the buffer is a window
        packed, off, nb, no = fq.pack_2bit(d, table + 1000, None, table_base=1000)
        assert np.array_equal(packed.cpu().numpy(), oracle.pack_2bit(data, host)[0])
    with pytest.raises(ValueError):
        fq.pack_2bit(d, table + len(data))


@pytest.mark.gpu
def test_length_filter_and_scan_match_oracle(fq, oracle):
    import torch
    from fastqandfurious_b200 import consume
    data = fqgen.variable_records_np(3000, 5, 'multiline').tobytes()
    d, table = _parsed(fq, data)
    host = table.cpu().numpy()
    for field in (0, 1, 2):
        lens = oracle.field_lengths(host, field)
        for lo, hi in ((0, 2 ** 62), (int(np.median(lens)), 2 ** 62), (0, int(np.median(lens))), (10 ** 9, 2 ** 62),
                       (int(lens.min()), int(lens.min()))):
            got = fq.select_by_length(table, lo, hi, field=field).cpu().numpy()
            assert np.array_equal(got, oracle.select_by_length(host, field, lo, hi)), (field, lo, hi)
    # prefix sums: block edges (2048 items per block), negative values, empty input
    rng = np.random.default_rng(3)
    for n in (0, 1, 2047, 2048, 2049, 70001, 600000):
        v = rng.integers(-5, 1000, size=n, dtype=np.int64)
        got = consume.exclusive_scan(torch.from_numpy(v).cuda()).cpu().numpy()
        want = np.zeros(n + 1, dtype=np.int64)
        want[1:] = np.cumsum(v)
        assert np.array_equal(got, want), n
    # a chunk of a stream: table positions are stream offsets, the buffer starts at table_base
    base = int(host[100, 0])
    sub = d[base:]
    sel = torch.arange(100, 200, dtype=torch.int64, device='cuda')
    out, off = fq.gather_fields(sub, table, 'sequence', sel, table_base=base)
    wout, woff = oracle.gather_fields(data, host, 1, list(range(100, 200)))
    assert np.array_equal(out.cpu().numpy(), wout) and np.array_equal(off.cpu().numpy(), woff)
    with pytest.raises(IndexError):
        fq.field_lengths(table, 'sequence', torch.tensor([len(host)], dtype=torch.int64, device='cuda'))
    with pytest.raises(ValueError):
        fq.gather_fields(sub, table, 'sequence', torch.tensor([0], dtype=torch.int64, device='cuda'), table_base=base)


@pytest.mark.gpu
def test_index_replay_end_to_end(fq, oracle, tmp_path):
    """Build the index on the GPU, write it like the reference's demo does, read it back and replay it on the
    device: same (header, sequence, quality) triples as the reference's readfastq_iter + entryfunc."""
    import torch
    data = fqgen.variable_records_np(500, 9, 'illumina').tobytes()
    table = fq.readfastq_table(io.BytesIO(data))
    path = tmp_path / 'reads.idx'
    with open(path, 'wb') as fh:
        fq.write_index(table, fh)
    with open(path, 'rb') as fh:
        idx = fq.read_index(fh, 'cuda')
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    fields = [fq.gather_fields(d, idx, f) for f in ('header', 'sequence', 'quality')]
    host = [(o.cpu().numpy().tobytes(), off.cpu().numpy()) for o, off in fields]
    ref = oracle.reference()
    if ref is not None:
        want = list(ref[0].readfastq_iter(io.BytesIO(data), 2 ** 16, entryfunc=ref[0].entryfunc, entrypos=ref[1].entrypos))
    else:
        t, _, _ = oracle.readfastq(data)
        want = [(data[r[0] + 1:r[1]], data[r[2]:r[3]], data[r[4]:r[5]]) for r in t]
    assert len(want) == len(idx)
    for k, trip in enumerate(want):
        got = tuple(host[f][0][host[f][1][k]:host[f][1][k + 1]] for f in range(3))
        assert got == tuple(trip), k
