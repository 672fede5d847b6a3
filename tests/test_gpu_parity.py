"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same bytes."""
import base64
import io
import json
import os
import random
from array import array

import numpy as np
import pytest

import fqgen

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


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
    """uint8 CUDA tensor holding `data`, starting `offset` bytes into an allocation (alignment tests)."""
    import torch
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    t = torch.empty(len(a) + offset + 32, dtype=torch.uint8, device='cuda')
    t.fill_(10 if offset else 0x40)  # hostile neighbours: newlines / '@' around the buffer
    if len(a):
        t[offset:offset + len(a)].copy_(torch.from_numpy(a.copy()))
    return t[offset:offset + len(a)]


def _check(fq, oracle, data, sentinel, goff, label='', **kw):
    if kw.get('force_general') and 'spec' not in kw:
        # the general path three times: exact resolution only, then with the speculative pass in front -- as one CTA per
        # chunk and as one warp per chunk (the default, the result returned)
        _check(fq, oracle, data, sentinel, goff, label, spec=False, **kw)
        _check(fq, oracle, data, sentinel, goff, label, spec='v1', **kw)
        kw['spec'] = True
    blob = (b'\n' if sentinel else b'') + bytes(data)
    want, st, tail, resume = oracle.parse_chain(blob, 0, goff)
    res = fq.parse_buffer(_dev(data, kw.pop('offset', 0)), sentinel=bool(sentinel), goff=goff, **kw)
    if kw.get('spec') is False:
        assert not res.spec
    got = res.table.cpu().numpy()
    ctx = (label, bytes(data)[:200], sentinel, goff, kw, res.path)
    assert res.n == len(want), ctx
    assert np.array_equal(got, want), ctx
    assert res.tail_status == st, ctx
    assert list(res.tail_pos) == tail.tolist(), ctx
    assert res.resume_offset == resume, ctx
    return res


@pytest.mark.parametrize('seed', range(4))
def test_corpus_auto_path(fq, oracle, seed):
    for data in fqgen.corpus(3000 + seed, 150):
        for sentinel in (1, 0):
            _check(fq, oracle, data, sentinel, -1)


@pytest.mark.parametrize('seed', range(4))
def test_corpus_general_path(fq, oracle, seed):
    for data in fqgen.corpus(4000 + seed, 150):
        for sentinel in (1, 0):
            res = _check(fq, oracle, data, sentinel, 5, force_general=True)
            assert res.path == 2


def test_clean_input_takes_fast_path(fq, oracle):
    rng = random.Random(11)
    for _ in range(40):
        data = fqgen.fastq_bytes(rng, rng.randint(1, 400), read_len=(20, 200), header_len=(5, 40), long_plus=0.3,
                                 trailing_newlines=rng.randint(0, 3), at_plus_bias=0.3)
        for cfg in range(6):
            res = _check(fq, oracle, data, 1, -1, cfg=cfg, offset=rng.randrange(16))
            assert res.path == 1


@pytest.mark.parametrize('name', ['test.fq', 'test_longqualityheader.fq', 'test_multiline.fq'])
def test_reference_data_files(fq, oracle, name):
    data = open(os.path.join(GOLD, name), 'rb').read()
    want, err, _ = oracle.readfastq(data)
    assert err == 0
    got = fq.readfastq_table(io.BytesIO(data), 100)
    assert np.array_equal(got, want)
    for chunk in (1, 64, 300, 1 << 20):
        assert np.array_equal(fq.readfastq_table(io.BytesIO(data), 1, device_chunk=chunk), want)
    # the reference's own call pattern, per record, through the plugin slots
    rows = [list(p) for p in fq.readfastq_iter(io.BytesIO(data), 600, entryfunc=fq.entryfunc_abspos)]
    assert rows == want.tolist()
    ents = list(fq.readfastq_iter(io.BytesIO(data), 700, entryfunc=fq.entryfunc_namedtuple))
    for e, r in zip(ents, want):
        assert (e.header, e.sequence, e.quality) == (data[r[0] + 1:r[1]], data[r[2]:r[3]], data[r[4]:r[5]])


def test_streams_vs_reference_golden(fq):
    """Every stream the unmodified reference was run on (tests/golden/make_golden.py)."""
    cases = json.load(open(os.path.join(GOLD, 'readfastq_kat.json')))
    n = 0
    for case in cases[:160]:
        data = base64.b64decode(case['data'])
        res = next((r for r in case['res'].values() if r is not None), None)
        if res is None:
            continue
        for chunk in (1 << 20, 37):
            try:
                rows = fq.readfastq_table(io.BytesIO(data), 1, device_chunk=chunk).tolist()
                err = None
            except ValueError as e:
                rows, err = e.rows.tolist(), str(e)
            assert rows == res['rows'], (case['name'], chunk)
            if res['error'] == 'LOOP':
                assert err is not None and err.startswith('Entry is invalid at byte')
            else:
                assert err == res['error'], (case['name'], chunk)
            n += 1
    assert n > 200


def test_entrypos_plugin_vs_reference_golden(fq):
    """The per-record plugin slot: DeviceEntryPos against the C extension's answers."""
    kat = json.load(open(os.path.join(GOLD, 'entrypos_kat.json')))
    rng = random.Random(3)
    ep = fq.DeviceEntryPos()
    for case in rng.sample(kat, 250):
        blob = base64.b64decode(case['blob'])
        pos = array('q', [-7] * 6)
        st = ep(blob, case['offset'], pos)
        assert [st, list(pos)] == case['c'], (blob, case['offset'])


def test_device_entrypos_is_stateless_for_the_caller(fq, oracle):
    """A bytearray rewritten in place between calls (same object, same length, any byte) is answered like a
    fresh call: the reference's entrypos keeps no state (src/_fastqandfurious.c:32,150-151)."""
    rng = random.Random(5)
    data = b'\n' + fqgen.fastq_bytes(rng, 300, read_len=(20, 120), at_plus_bias=0.3)
    ba = bytearray(data)
    ep = fq.DeviceEntryPos(min_window=4096, max_window=1 << 16)
    pos, want = array('q', [-1] * 6), array('q', [-1] * 6)
    off = 0
    for step in range(120):
        st = ep(ba, off, pos)
        assert (st, list(pos)) == (oracle.entrypos(bytes(ba), off, want), list(want)), step
        if step % 3 == 0:  # damage a byte somewhere ahead, then ask again at the same offset
            i = rng.randrange(off + 1, len(ba))
            ba[i] = rng.choice(b'\n@+A')
            st = ep(ba, off, pos)
            assert (st, list(pos)) == (oracle.entrypos(bytes(ba), off, want), list(want)), (step, i)
        if st != fq.COMPLETE:
            break
        off = pos[5] - 1
    # the whole chain through small windows equals the oracle's chain
    big = b'\n' + fqgen.variable_records_np(400, 4, 'multiline').tobytes()
    ep = fq.DeviceEntryPos(min_window=2048, max_window=1 << 15)
    off, rows = 0, []
    while ep(big, off, pos) == fq.COMPLETE:
        off = pos[5] - 1
        rows.append(list(pos))
    wrows, wst, wtail, _ = oracle.parse_chain(big, 0, 0)
    assert rows == wrows.tolist()


def test_reference_loop_over_device_entrypos(fq, oracle):
    """The reference's unmodified loop shape (entrypos then entryfunc, offset = pos5 - 1)."""
    data = fqgen.variable_records_np(200, 9, 'multiline').tobytes()
    want, err, _ = oracle.readfastq(data)
    buf = b'\n' + data
    ep = fq.DeviceEntryPos()
    pos = array('q', [-1] * 6)
    off, rows = 0, []
    while ep(buf, off, pos) == fq.COMPLETE:
        off = pos[5] - 1
        rows.append([p - 1 for p in pos])
    assert rows == want[:len(rows)].tolist() and len(want) - len(rows) <= 1


@pytest.mark.parametrize('kind,nrec', [('illumina', 20000), ('ont', 300), ('multiline', 15000)])
def test_config_shapes_vs_oracle(fq, oracle, kind, nrec):
    """BASELINE.json configs 3-5 at oracle-checkable size, offsets and decoded qualities."""
    data = fqgen.variable_records_np(nrec, 17, kind).tobytes()
    res = _check(fq, oracle, data, 1, -1, decode_quality=True)
    assert res.path == (2 if kind == 'multiline' else 1)
    if kind == 'multiline':  # clean wrapped records: the speculative pass answers, the exact path is not needed
        assert res.spec
    r2 = _check(fq, oracle, data, 1, -1, force_general=True)
    assert r2.spec == (kind != 'ont')  # long reads leave the window: declined, resolved exactly
    got = res.table.cpu().numpy()
    q = res.qual.cpu().numpy()
    want_q = oracle.decode_quals(data, got)
    got_q = np.concatenate([q[r[4]:r[5]] for r in got])
    assert np.array_equal(got_q, want_q)
    tab = fq.readfastq_table(io.BytesIO(data), device_chunk=1 << 20)
    assert len(tab) == nrec and np.array_equal(tab, oracle.readfastq(data)[0])


def test_speculative_general_pass_on_damaged_four_line_input(fq, oracle):
    """A 4-line file with a few damaged records (the fast path declines the whole buffer): the speculative pass
    follows the reference through every resynchronisation; INVALID entries and tiny lines make it decline -- same
    table either way."""
    rng = random.Random(21)
    n_spec = 0
    for trial in range(24):
        base = fqgen.variable_records_np(4000, 40 + trial, 'illumina').tobytes()
        data = fqgen.mutate(rng, base, rng.randint(1, 3)) if trial % 4 else base[:-1 - rng.randrange(300)]
        res = _check(fq, oracle, data, 1, -1, force_general=True)
        auto = _check(fq, oracle, data, 1, -1)
        assert auto.n == res.n
        n_spec += bool(res.spec)
    assert n_spec >= 12
    # dense lines do not fit a window: declined, exact
    tiny = fqgen.fastq_bytes(rng, 30000, read_len=(1, 3), header_len=(0, 1))
    res = _check(fq, oracle, tiny, 1, -1, force_general=True)
    assert not res.spec


def test_fixed150_vs_oracle_and_closed_form(fq, oracle):
    """BASELINE.json config 2 shape at 64 MiB against the oracle, every kernel configuration."""
    n = (64 << 20) // 337
    d = fq.synth_fixed(n)
    host = d.cpu().numpy()
    assert np.array_equal(host[:337 * 1000], fqgen.fixed_records_np(1000))
    want, st, tail, resume = oracle.parse_chain(np.concatenate([np.array([10], np.uint8), host]), 0, -1)
    for cfg in range(6):
        res = fq.parse_buffer(d, cfg=cfg, decode_quality=(cfg % 3 == 0))
        assert res.path == 1 and res.n == len(want) == n - 1  # the last record needs the EOF rule
        assert np.array_equal(res.table.cpu().numpy(), want)
        assert (res.tail_status, list(res.tail_pos), res.resume_offset) == (st, tail.tolist(), resume)
        if res.qual is not None:
            q = res.qual.cpu().numpy().reshape(n, 337)[:, 186:336]
            assert np.array_equal(q[:-1], (host.reshape(n, 337)[:-1, 186:336].astype(np.int16) - 33).astype(np.int8))
    res = fq.parse_buffer(d, force_general=True)
    assert res.path == 2 and np.array_equal(res.table.cpu().numpy(), want)


def test_full_size_config2_properties(fq):
    """BASELINE.json config 2 at full size (1 GiB): closed-form offsets and a checksum."""
    import torch
    n = (1 << 30) // 337
    d = fq.synth_fixed(n)
    res = fq.parse_buffer(d, decode_quality=True)
    assert res.path == 1 and res.n == n - 1 and res.tail_status == fq.MISSING_QUAL_END
    k = torch.arange(n - 1, device='cuda', dtype=torch.int64) * 337
    want = torch.stack([k, k + 32, k + 33, k + 183, k + 186, k + 336], dim=1)
    assert torch.equal(res.table, want)
    q = res.qual.view(n, 337)[:-1, 186:336]
    assert torch.equal(q, (d.view(n, 337)[:-1, 186:336].to(torch.int16) - 33).to(torch.int8))
    tab = fq.readfastq_table(_TensorReader(d), device_chunk=1 << 28)
    assert len(tab) == n and np.array_equal(tab[:, 0], np.arange(n, dtype=np.int64) * 337)
    assert np.array_equal(tab[:, 5] - tab[:, 4], np.full(n, 150))


class _TensorReader:
    """Minimal .read(n) object over a CUDA tensor's host copy (what readfastq_iter needs, :30-36)."""

    def __init__(self, t):
        self.b = t.cpu().numpy().tobytes()
        self.o = 0

    def read(self, n):
        out = self.b[self.o:self.o + n]
        self.o += len(out)
        return out


def test_unaligned_and_hostile_neighbours(fq, oracle):
    rng = random.Random(5)
    data = fqgen.fastq_bytes(rng, 300, read_len=(30, 90), header_len=(3, 20), trailing_newlines=1, at_plus_bias=0.4)
    for off in range(16):
        for sentinel in (1, 0):
            res = _check(fq, oracle, data, sentinel, 1000, offset=off, decode_quality=True)
            got = res.table.cpu().numpy() - 1000 - sentinel
            q = res.qual.cpu().numpy()
            want_q = oracle.decode_quals(data, got)
            assert np.array_equal(np.concatenate([q[r[4]:r[5]] for r in got]), want_q)


def test_empty_and_tiny(fq, oracle):
    for data in (b'', b'\n', b'@', b'@\n', b'\n@', b'@a\nA\n+\nI', b'@a\nA\n+\nI\n', b'@a\nA\n+\nI\n\n', b'+', b'\n\n\n'):
        for sentinel in (1, 0):
            if ((b'\n' if sentinel else b'') + data).endswith(b'\n@'):
                blob = (b'\n' if sentinel else b'') + data
                res = fq.parse_buffer(_dev(data), sentinel=bool(sentinel), goff=0)
                assert res.tail_status == fq.MISSING_SEQHEADER_END  # the reference's UB case: defined here
                continue
            _check(fq, oracle, data, sentinel, 0)
            _check(fq, oracle, data, sentinel, 0, force_general=True)


def test_arrayadd(fq, oracle):
    import torch
    rng = np.random.default_rng(1)
    for n in (0, 1, 15, 16, 17, 1000, 1 << 20):
        for off in (0, 3):
            a = rng.integers(-128, 128, n + off).astype(np.int8)
            for v in (-33, 127, -128, 1000):
                t = torch.from_numpy(a.copy()).cuda()
                fq.device.arrayadd_b_(t[off:], v)
                w = a.copy()
                oracle.arrayadd_b(w[off:], ((v + 2 ** 15) % 2 ** 16) - 2 ** 15)
                assert np.array_equal(t.cpu().numpy(), w)
    q = rng.integers(-2 ** 62, 2 ** 62, 1001).astype(np.int64)
    t = torch.from_numpy(q.copy()).cuda()
    fq.device.arrayadd_q_(t, -12345678901)
    w = q.copy()
    oracle.arrayadd_q(w, -12345678901)
    assert np.array_equal(t.cpu().numpy(), w)
    for case in json.load(open(os.path.join(GOLD, 'arrayadd_kat.json'))):
        a = array(case['kind'], case['in'])
        (fq.arrayadd_b if case['kind'] == 'b' else fq.arrayadd_q)(a, case['value'])
        assert list(a) == case['out']
    with pytest.raises(ValueError, match='format type q'):
        fq.arrayadd_q(array('b', [1]), 1)
    # the module-level functions on CUDA tensors: in place, None returned like the reference (any number of elements)
    for n in (0, 1, 2, 4097):
        a = rng.integers(-128, 128, n).astype(np.int8)
        t = torch.from_numpy(a.copy()).cuda()
        assert fq.arrayadd_b(t, -33) is None
        w = a.copy()
        oracle.arrayadd_b(w, -33)
        assert np.array_equal(t.cpu().numpy(), w)
        q = rng.integers(-2 ** 40, 2 ** 40, n).astype(np.int64)
        t = torch.from_numpy(q.copy()).cuda()
        assert fq.arrayadd_q(t, 7) is None
        assert np.array_equal(t.cpu().numpy(), q + 7)
    z = torch.tensor([33], dtype=torch.int8, device='cuda')  # a one-element result of 0 is still None
    assert fq.arrayadd_b(z, -33) is None and int(z.item()) == 0


@pytest.mark.parametrize('kind,gib,nrec', [('illumina', 64, 120000), ('ont', 8, 3000), ('multiline', 8, 60000)])
def test_full_size_configs_by_periodicity(fq, oracle, kind, gib, nrec):
    """BASELINE.json configs 3-5 at full size (64 / 8 / 8 GiB on one GPU): the buffer is a record-aligned
    block, checked against the oracle, repeated to full size, so the offsets must be periodic."""
    import torch
    base = fqgen.variable_records_np(nrec, 23, kind)
    want = oracle.readfastq(base.tobytes())[0]
    n0, blen = len(want), len(base)
    assert n0 == nrec
    reps = (gib << 30) // blen
    d = torch.from_numpy(base).cuda().repeat(reps)
    try:
        res = fq.parse_buffer(d, cap=n0 * reps + 64)
        assert res.path == (2 if kind == 'multiline' else 1)
        assert res.n == n0 * reps - 1 and res.tail_status == fq.MISSING_QUAL_END  # last record: EOF rule
        w = torch.from_numpy(want).cuda()
        full = res.table[:(reps - 1) * n0].view(reps - 1, n0, 6)
        shift = (torch.arange(reps - 1, device='cuda', dtype=torch.int64) * blen).view(-1, 1, 1)
        assert torch.equal(full - shift, w.unsqueeze(0).expand(reps - 1, n0, 6))
        last = res.table[(reps - 1) * n0:] - (reps - 1) * blen
        assert torch.equal(last, w[:n0 - 1])
        # sortedness + checksum of the whole table as size-independent properties
        assert bool((res.table[1:, 0] > res.table[:-1, 5]).all())
    finally:
        del d
        fq.device._ws_cache.clear()
        torch.cuda.empty_cache()


def test_fused_decode_mirror_alignment_and_bounds(fq, oracle):
    """The Phred mirror written by the scan itself (mirror congruent to the buffer modulo 16: whole-tile vector
    stores, edge tiles byte by byte) and by the span kernel (any other alignment): the quality spans hold
    byte - 33 and nothing outside the mirror is touched."""
    import torch
    rng = random.Random(11)
    small = fqgen.fastq_bytes(rng, 40, read_len=(30, 90), header_len=(3, 20), trailing_newlines=1)
    big = fqgen.fixed_records_np(700).tobytes()  # 236 KB: interior tiles and both edges
    for data in (small, big, big[:-1], big[:100003]):
        want, _, _ = oracle.readfastq(data)
        for off_b, off_q in ((0, 0), (5, 5), (13, 13 + 16), (0, 3), (7, 2), (16, 0)):
            d = _dev(data, off_b)
            guard = 64
            raw = torch.full((len(data) + off_q + 2 * guard,), 0x5a, dtype=torch.int8, device='cuda')
            qual = raw[guard + off_q:guard + off_q + len(data)]
            res = fq.parse_buffer(d, decode_quality=True, qual=qual)
            got = res.table.cpu().numpy()
            assert np.array_equal(got, want[:len(got)]) and len(want) - len(got) <= 1
            host = raw.cpu().numpy()
            assert (host[:guard + off_q] == 0x5a).all() and (host[guard + off_q + len(data):] == 0x5a).all(), (off_b, off_q)
            q = host[guard + off_q:guard + off_q + len(data)]
            assert np.array_equal(np.concatenate([q[r[4]:r[5]] for r in got]), oracle.decode_quals(data, got)), (off_b, off_q)


@pytest.mark.parametrize('frac', [0.05, 0.5])
def test_speculative_pass_with_many_false_candidates(fq, oracle, frac):
    """Wrapped records (about 200 chunks of the warp-per-chunk pass) in which a share of the QUALITY lines begin with
    '@' or '+' -- false candidates the chunks' speculated entries must not start from, and '+' lines that are not the
    record's.  Whether the speculative pass answers or declines, the table is the reference's; with few false
    candidates it must answer."""
    data = bytearray(fqgen.variable_records_np(30000, 91, 'multiline').tobytes())
    tab = oracle.readfastq(bytes(data))[0]
    rng = np.random.default_rng(int(frac * 100))
    nl = np.frombuffer(bytes(data), dtype=np.uint8) == 10
    for r in tab[rng.random(len(tab)) < frac]:
        starts = [int(r[4])] + [int(i) + 1 for i in np.flatnonzero(nl[r[4]:r[5]]) + r[4] if i + 1 < r[5]]
        for s in starts:
            if rng.random() < 0.6:
                data[s] = 0x40 if rng.random() < 0.7 else 0x2b
    data = bytes(data)
    assert np.array_equal(oracle.readfastq(data)[0], tab)  # quality bytes do not move the reference's chain
    answered = {}
    for spec in ('v1', True):
        res = _check(fq, oracle, data, 1, -1, spec=spec)
        assert res.path == 2
        answered[spec] = bool(res.spec)
    if frac <= 0.05:
        assert answered[True], answered


def test_host_parser_adapts_the_scan_geometry_to_dense_lines(fq, oracle):
    """HostParser over several chunks of wrapped records: from the second chunk on the scan runs the geometry the
    first chunk's line density asks for (device.geometry_for_density) -- same table as the reference's loop."""
    import torch
    data = fqgen.variable_records_np(24000, 5, 'multiline').tobytes()
    want = oracle.readfastq(data)[0]
    hp = fq.device.HostParser('cuda', chunk_bytes=3 << 20)
    got = hp.parse(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    assert np.array_equal(got, want)
    assert hp.stats['device_calls'] >= 4
    assert fq.device.geometry_for_density(0, 22_000_000, 1 << 30) == 2      # wrapped records: 47 bytes per line
    assert fq.device.geometry_for_density(0, 12_700_000, 1 << 30) == 0      # 150 bp, four lines per record
    assert fq.device.geometry_for_density(3, 22_000_000, 1 << 30) == 3      # the caller's choice stands
