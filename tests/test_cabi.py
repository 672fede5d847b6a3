"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/fqb200.h
declares, and the Python mirror of the reference interface is importable (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'fqb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(fqb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from fastqandfurious_b200 import _lib
    declared = _declared_symbols()
    assert set(declared) == set(_lib.SYMBOLS), declared
    L = ctypes.CDLL(_lib.LIBPATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().fqb_version().startswith(b'fqb200')
    # pure host arithmetic: workspace size grows with the buffer and with the line budget
    w0 = _lib.lib().fqb_workspace_bytes(1 << 20, 0, 0)
    w1 = _lib.lib().fqb_workspace_bytes(1 << 30, 0, 0)
    w2 = _lib.lib().fqb_workspace_bytes(1 << 20, 1 << 20, 0)
    assert 0 < w0 < w1 and w2 > w0 + 30 * (1 << 20)


def test_result_struct_matches_header():
    from fastqandfurious_b200 import _lib
    assert ctypes.sizeof(_lib.FqbResult) == 128
    text = open(os.path.join(ROOT, 'include', 'fqb200.h')).read()
    for name, val in (('FQB_INVALID', -1), ('FQB_COMPLETE', 6), ('FQB_MISSING_QUALHEADER_END', 7),
                      ('FQB_MISSING_QUAL_END', 5), ('FQB_MISSING_SEQ_END', 3)):
        m = re.search(r'#define %s \(?(-?\d+)\)?' % name, text)
        assert m and int(m.group(1)) == val


def test_module_surface_mirrors_reference():
    import fastqandfurious_b200 as fq
    for name in ('readfastq_iter', 'entrypos', 'entryfunc', 'entryfunc_namedtuple', 'entryfunc_abspos', 'Entry',
                 'read', 'arrayadd_b', 'arrayadd_q', 'INVALID', 'MISSING_SEQHEADER_BEGIN', 'MISSING_SEQHEADER_END',
                 'MISSING_SEQ_BEG', 'MISSING_SEQ_END', 'MISSING_QUAL_BEGIN', 'MISSING_QUAL_END', 'COMPLETE',
                 'MISSING_QUALHEADER_END'):
        assert hasattr(fq, name), name
    ce = fq._fastqandfurious
    for name in ('entrypos', 'arrayadd_b', 'arrayadd_q', 'INVALID', 'POS_HEAD_BEG', 'POS_HEAD_END', 'POS_SEQ_BEG',
                 'POS_SEQ_END', 'POS_QUAL_BEG', 'POS_QUAL_END', 'COMPLETE', 'MISSING_QUALHEADER_END'):
        assert hasattr(ce, name), name
    assert (fq.INVALID, fq.COMPLETE, fq.MISSING_QUALHEADER_END) == (-1, 6, 7)


def test_reference_loop_with_host_plugin_matches_golden(oracle):
    """readfastq_iter with a caller-supplied entrypos runs the reference's per-record loop: host logic,
    checked with the oracle's entrypos as the plugin."""
    import io
    from array import array
    import fastqandfurious_b200 as fq
    gold = os.path.join(ROOT, 'tests', 'golden')
    data = open(os.path.join(gold, 'test_multiline.fq'), 'rb').read()
    for fb in (100, 200, 600, 700):
        rows = [list(p) for p in fq.readfastq_iter(io.BytesIO(data), fb, entryfunc=fq.entryfunc_abspos,
                                                   entrypos=oracle.entrypos)]
        assert rows == [[0, 30, 31, 67, 99, 135], [136, 166, 167, 203, 206, 242], [243, 272, 273, 309, 312, 348],
                        [349, 379, 380, 417, 420, 457]]
    with pytest.raises(ValueError, match='Incomplete final quality string at byte'):
        list(fq.readfastq_iter(io.BytesIO(b'@r1\nACGT\n+\nIIII\n@r2\nGG\n+\nII'), 50, entrypos=oracle.entrypos))
    hdr, seq, qual = next(fq.readfastq_iter(io.BytesIO(b'@r1\nACGT\n+\nIIII\n'), 50, entrypos=oracle.entrypos))
    assert (hdr, seq, qual) == (b'r1', b'ACGT', b'IIII')
    pos = array('q', [1, 2, 3, 4, 5, 6])
    assert fq.entryfunc_abspos(b'', pos, 10) is pos and list(pos) == [11, 12, 13, 14, 15, 16]


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import io
    import fastqandfurious_b200 as fq
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        list(fq.readfastq_iter(io.BytesIO(b'@r1\nACGT\n+\nIIII\n'), 50))


def test_automagic_open_by_extension(tmp_path):
    """Host I/O mirror of src/fastqandfurious.py:290-334: the opener is chosen by the file extension."""
    import bz2
    import gzip
    import lzma
    import fastqandfurious_b200 as fq
    data = open(os.path.join(ROOT, 'tests', 'golden', 'test.fq'), 'rb').read()
    for ext, mod in (('fq', None), ('fq.gz', gzip), ('fq.gzip', gzip), ('fq.bz2', bz2), ('fq.lzma', lzma), ('noext', None)):
        path = str(tmp_path / ('reads.' + ext)) if ext != 'noext' else str(tmp_path / 'reads')
        with (mod.open(path, 'wb') if mod else open(path, 'wb')) as fh:
            fh.write(data)
        with fq.automagic_open(path) as fh:
            assert fh.read() == data
    # a caller-supplied table is honoured (namespace object instead of a module name)
    import io
    path = str(tmp_path / 'reads.fq.gz')
    with fq.automagic_open(path, openers={'gz': (io, 'open', ('rb',))}) as fh:
        assert fh.read(2) == b'\x1f\x8b'


def test_fast_source_fills_staging_buffers_like_readinto(tmp_path, monkeypatch):
    """Host logic of readfastq_table's reader thread: a regular file is copied by several threads (os.preadv), any
    other object through readinto() / read(); same bytes, same short count at the end of the stream, and the
    object's own position stays in step."""
    import io

    import numpy as np
    import __graft_entry__  # noqa: F401  (puts the package on sys.path)
    from fastqandfurious_b200 import api
    monkeypatch.setattr(api._FastSource, 'MIN_SLICE', 1000)
    data = np.random.default_rng(5).integers(0, 256, 100_003, dtype=np.uint8).tobytes()
    path = tmp_path / 'blob.bin'
    path.write_bytes(data)

    class ReadOnly:  # no readinto, no fileno: the generic route
        def __init__(self, d):
            self.d, self.p = d, 0

        def read(self, n):
            r = self.d[self.p:self.p + min(n, 777)]
            self.p += len(r)
            return r

    for make, kind in ((lambda: io.BytesIO(data), 'generic'), (lambda: open(path, 'rb'), 'file'),
                       (lambda: ReadOnly(data), 'generic')):
        for threads in (1, 5):
            fh = make()
            lead = fh.read(13)
            assert lead == data[:13]
            src = api._FastSource(fh, threads=threads)
            assert src.kind == kind
            dst = np.zeros(30_000, dtype=np.uint8)
            got = bytearray(lead)
            while True:
                n = src.readinto(memoryview(dst))
                got += dst[:n].tobytes()
                if hasattr(fh, 'tell'):
                    assert fh.tell() == len(got)
                if n < len(dst):
                    break
            assert bytes(got) == data
            assert src.readinto(memoryview(dst)) == 0
            src.close()
            if hasattr(fh, 'close'):
                fh.close()


def test_entry_points_reject_bad_arguments_before_touching_the_device():
    """Every C entry point validates its scalar and pointer arguments first and answers cudaErrorInvalidValue (1):
    no kernel, no CUDA call -- so this runs without a GPU.  (What the C extension leaves unchecked upstream --
    posbuffer length, src/_fastqandfurious.c:38-43 -- is checked here.)"""
    import __graft_entry__ as g
    g.build()
    from fastqandfurious_b200 import _lib
    L = _lib.lib()
    INVALID_VALUE = 1
    nul = None
    fake = ctypes.c_void_p(256)  # never dereferenced: the calls below fail on an earlier argument
    # fqb_parse: negative capacity, missing result header, rows wanted but no table
    assert L.fqb_parse(nul, 0, 1, -1, nul, -1, nul, 0, fake, fake, 1 << 20, 0, 0, nul) == INVALID_VALUE
    assert L.fqb_parse(nul, 0, 1, -1, nul, 0, nul, 0, nul, fake, 1 << 20, 0, 0, nul) == INVALID_VALUE
    assert L.fqb_parse(nul, 0, 1, -1, nul, 8, nul, 0, fake, fake, 1 << 20, 0, 0, nul) == INVALID_VALUE
    # negative length / missing or misaligned workspace
    assert L.fqb_parse(nul, -5, 1, -1, nul, 0, nul, 0, fake, fake, 1 << 20, 0, 0, nul) == INVALID_VALUE
    assert L.fqb_parse(nul, 0, 1, -1, nul, 0, nul, 0, fake, nul, 0, 0, 0, nul) == INVALID_VALUE
    assert L.fqb_parse(nul, 0, 1, -1, nul, 0, nul, 0, fake, ctypes.c_void_p(257), 1 << 20, 0, 0, nul) == INVALID_VALUE
    assert L.fqb_parse(nul, 64, 1, -1, nul, 0, nul, 0, fake, fake, 1 << 20, 0, 0, nul) == INVALID_VALUE  # bytes but no buffer
    # sharded entry points
    assert L.fqb_shard_scan(nul, 0, 0, 1, nul, fake, 1 << 20, 0, nul) == INVALID_VALUE
    assert L.fqb_shard_scan(nul, 10, 11, 1, fake, fake, 1 << 20, 0, nul) == INVALID_VALUE  # own_len > len
    assert L.fqb_shard_scan_publish(nul, 0, 0, 1, fake, nul, 3, 0, fake, 1 << 20, 0, nul) == INVALID_VALUE
    assert L.fqb_shard_scan_publish(nul, 0, 0, 1, fake, nul, 17, 1, fake, 1 << 20, 0, nul) == INVALID_VALUE
    assert L.fqb_shard_scan_decode(nul, 0, 0, 1, fake, nul, 0, 0, nul, -33, fake, 1 << 20, 0, nul) == INVALID_VALUE
    assert L.fqb_shard_emit(nul, 0, 0, 1, 1, 0, nul, nul, 0, fake, fake, 1 << 20, 0, nul) == INVALID_VALUE
    # consumers of the table
    assert L.fqb_field_lengths(nul, 0, nul, 0, 3, nul, nul, nul) == INVALID_VALUE          # no such field
    assert L.fqb_field_lengths(nul, 4, nul, 3, 1, fake, nul, nul) == INVALID_VALUE         # all rows: n_sel == n_rows
    assert L.fqb_gather_fields(nul, -1, 0, nul, 0, nul, 0, 1, nul, nul, 0, nul, nul) == INVALID_VALUE
    assert L.fqb_gather_fields(nul, 0, 0, nul, 0, nul, -2, 1, nul, nul, 0, nul, nul) == INVALID_VALUE
    assert L.fqb_pack_2bit(nul, -1, 0, nul, 0, nul, 0, nul, nul, nul, nul, nul, nul) == INVALID_VALUE
    assert L.fqb_pack_2bit(nul, 8, 0, fake, 2, nul, 2, nul, nul, nul, nul, nul, nul) == INVALID_VALUE  # bytes but no buffer
    assert L.fqb_exclusive_scan(nul, 4, fake, nul, 0, nul) != 0
    assert L.fqb_parse_fasta(nul, 0, 1, -1, nul, 0, fake, fake, 1 << 20, 0, 0, nul) == INVALID_VALUE   # max_lines < 1
    assert L.fqb_kernel_info(-1, nul, nul, nul, nul) == INVALID_VALUE
    assert L.fqb_kernel_info(99, nul, nul, nul, nul) == INVALID_VALUE
    # sizes are pure host arithmetic: multiples of 256, monotonic, FASTA needs more than the plain scan
    for n in (0, 1, 4096, 1 << 24):
        w = L.fqb_workspace_bytes(n, 0, 0)
        assert w % 256 == 0 and w >= L.fqb_workspace_bytes(max(n - 1, 0), 0, 0)
        assert L.fqb_workspace_bytes(n, 0, _lib.FLAG_DENSE) >= w
        assert L.fqb_fasta_workspace_bytes(n, n // 2 + 64, 0) > w
        assert L.fqb_scan_workspace_bytes(n) % 256 == 0


class _FakeStager:
    """Host stand-in of api._Stager for the CPU tests: keeps the uploaded bytes as the `pinned` snapshot."""

    def __init__(self):
        import torch
        self._torch = torch
        self.pinned = None
        self.uploads = []

    def upload(self, blob):
        import numpy as np
        self.pinned = self._torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy())
        self.uploads.append(len(blob))
        return self.pinned


def _host_entrypos(api, oracle, monkeypatch, max_window=1 << 26, min_window=None):
    """A DeviceEntryPos whose device chain is computed by the oracle (host logic under test: the chain cache)."""
    monkeypatch.setattr(api, '_device', lambda dev: 'cpu')
    import contextlib
    monkeypatch.setattr(api.torch.cuda, 'device', lambda dev: contextlib.nullcontext())
    ep = api.DeviceEntryPos(max_window=max_window, min_window=min_window)
    ep._stager = _FakeStager()

    def chain(d, off):
        blob = d.numpy().tobytes()
        rows, st, tail, resume = oracle.parse_chain(blob, 0, off)
        return rows, st, [int(p) + off if p >= 0 else -1 for p in tail]
    ep._device_chain = chain
    return ep


def _walk(ep, buf, api, entryfunc=None):
    """The reference's inner loop (src/fastqandfurious.py:251-255) over one buffer."""
    from array import array
    pos = array('q', [-1] * 6)
    offset, out = 0, []
    while True:
        st = ep(buf, offset, pos)
        if st != api.COMPLETE:
            return out, st, list(pos)
        out.append(list(pos))
        offset = pos[5] - 1


def test_chain_cache_never_changes_an_answer(oracle, monkeypatch):
    """DeviceEntryPos answers the reference's per-record calls from a chain cached per buffer.  The reference's
    entrypos is stateless (src/_fastqandfurious.c:32,150-151), so the cache must be invisible: mutable buffers
    rewritten in place (anywhere, not only at sampled positions), windows shorter than the buffer, calls off the
    chain and calls from several buffers in turn all get the answer of a fresh call."""
    import random
    from array import array
    import __graft_entry__  # noqa: F401
    from fastqandfurious_b200 import api
    import fqgen
    rng = random.Random(11)

    def truth(buf, offset):
        pos = array('q', [-1] * 6)
        st = oracle.entrypos(bytes(buf), offset, pos)
        return st, list(pos)

    data = b'\n' + fqgen.fastq_bytes(rng, 60, read_len=(5, 40), at_plus_bias=0.3)
    # (1) immutable bytes: the whole chain from one parse
    ep = _host_entrypos(api, oracle, monkeypatch)
    rows, st, tail = _walk(ep, data, api)
    want, wst, wtail, _ = oracle.parse_chain(data, 0, 0)
    assert rows == want.tolist() and st == wst and tail == wtail.tolist()
    assert len(ep._stager.uploads) == 1
    assert ep._buf is None and ep._id is None  # released once the tail has been handed out
    # (2) a bytearray rewritten in place BETWEEN the sampled positions of the old fingerprint (any single byte)
    ba = bytearray(data)
    ep = _host_entrypos(api, oracle, monkeypatch)
    pos = array('q', [-1] * 6)
    assert ep(ba, 0, pos) == api.COMPLETE
    first = list(pos)
    off1 = first[5] - 1
    assert ep(ba, off1, pos) == api.COMPLETE and [truth(ba, off1)] == [(api.COMPLETE, list(pos))]
    for _ in range(200):
        i = rng.randrange(1, len(ba))
        old = ba[i]
        ba[i] = rng.choice(b'\n@+AI')
        offset = rng.choice([0, off1, first[5] - 1, rng.randrange(len(ba))])
        st = ep(ba, offset, pos)
        assert (st, list(pos)) == truth(ba, offset), (i, offset)
        ba[i] = old
        st = ep(ba, offset, pos)
        assert (st, list(pos)) == truth(ba, offset), (i, offset)
    # a refill with different records of the same total length
    other = b'\n' + fqgen.fastq_bytes(random.Random(12), 200, read_len=(5, 40))
    ba2 = bytearray(data)
    assert ep(ba2, 0, pos) == api.COMPLETE
    ba2[:] = (other + b'\n' * len(ba2))[:len(ba2)]
    rows2, st2, tail2 = _walk(ep, ba2, api)
    want2, wst2, wtail2, _ = oracle.parse_chain(bytes(ba2), 0, 0)
    assert rows2 == want2.tolist() and st2 == wst2 and tail2 == wtail2.tolist()
    # (3) windows shorter than the buffer: the chain is continued window by window, windows grow while the caller
    #     follows the chain, a first record longer than the window doubles it
    big = b'\n' + fqgen.fastq_bytes(random.Random(13), 400, read_len=(20, 300), at_plus_bias=0.2)
    ep = _host_entrypos(api, oracle, monkeypatch, max_window=4096, min_window=64)
    rows, st, tail = _walk(ep, big, api)
    want, wst, wtail, _ = oracle.parse_chain(big, 0, 0)
    assert rows == want.tolist() and st == wst and tail == wtail.tolist()
    ups = ep._stager.uploads
    assert len(ups) > 3 and max(ups) <= 4096 and ups[0] < max(ups)
    # damaged input (INVALID / resynchronisation) through small windows
    for seed in range(30):
        r2 = random.Random(100 + seed)
        bad = b'\n' + fqgen.mutate(r2, fqgen.fastq_bytes(r2, 30, read_len=(5, 60), long_plus=0.4), 3)
        ep = _host_entrypos(api, oracle, monkeypatch, max_window=1024, min_window=64)
        rows, st, tail = _walk(ep, bad, api)
        want, wst, wtail, _ = oracle.parse_chain(bad, 0, 0)
        assert rows == want.tolist() and st == wst and tail == wtail.tolist(), seed
    # (4) random access: every answer equals a fresh call, and an off-chain call uploads a bounded window
    ep = _host_entrypos(api, oracle, monkeypatch, max_window=4096, min_window=256)
    for _ in range(100):
        offset = rng.randrange(len(big))
        st = ep(big, offset, pos)
        assert (st, list(pos)) == truth(big, offset)
    assert max(ep._stager.uploads) <= max(4096, len(big) // 2)
    # (5) two buffers in turn
    ep = _host_entrypos(api, oracle, monkeypatch)
    a, b = data, big
    pa, pb = array('q', [-1] * 6), array('q', [-1] * 6)
    oa = ob = 0
    for _ in range(20):
        assert (ep(a, oa, pa), list(pa)) == truth(a, oa)
        assert (ep(b, ob, pb), list(pb)) == truth(b, ob)
        oa, ob = pa[5] - 1, pb[5] - 1
    ep.reset()
    assert ep._id is None and ep._window == ep.MIN_WINDOW


def test_globaloffset_is_honoured_as_base_offset(oracle):
    """`globaloffset` (src/fastqandfurious.py:198-203) is accepted and then overwritten with -1 upstream (:242); here
    it is the absolute position of the stream's first byte: the default 0 reproduces the reference, any other value
    shifts every absolute position and the byte quoted in error messages."""
    import io
    import __graft_entry__  # noqa: F401
    import fastqandfurious_b200 as fq
    data = open(os.path.join(ROOT, 'tests', 'golden', 'test.fq'), 'rb').read()
    want, err, _ = oracle.readfastq(data)
    for base in (0, 1000, -7):
        rows = [list(p) for p in fq.readfastq_iter(io.BytesIO(data), 200, entryfunc=fq.entryfunc_abspos,
                                                   entrypos=oracle.entrypos, globaloffset=base)]
        assert rows == (want + base).tolist()
    with pytest.raises(ValueError, match='Incomplete entry at byte %d' % (14 + 500)):
        list(fq.readfastq_iter(io.BytesIO(b'@r1\nACGT\n+\nIIII\n@r2\nAC'), 64, entrypos=oracle.entrypos, globaloffset=500))


def test_readfastq_iter_replays_chunk_tables_through_entryfunc(oracle, monkeypatch):
    """Host logic of the drop-in readfastq_iter: the per-chunk offset tables (here produced by the oracle instead of
    the device) are replayed record by record through the caller's entryfunc exactly as the reference's loop hands
    them over: a fresh 6-item array('q') per record, positions relative to the chunk, globaloffset of the chunk."""
    import io
    from array import array

    import numpy as np
    import __graft_entry__  # noqa: F401
    from fastqandfurious_b200 import api
    data = open(os.path.join(ROOT, 'tests', 'golden', 'test_longqualityheader.fq'), 'rb').read()
    blob = b'\n' + data
    table, err, _ = oracle.readfastq(data)
    assert err == 0 and len(table) == 4
    rel = table + 1  # chunk-relative positions for globaloffset -1

    def fake_chunks(fh, fbufsize, device_chunk, dev, decode_quality=False, stats=None, base=0):
        yield blob, rel[:3], None, -1
        yield blob, rel[3:], None, -1
    monkeypatch.setattr(api, '_chunks', fake_chunks)
    monkeypatch.setattr(api, '_device', lambda dev: 'cpu')
    seen = []

    def spy(buf, pos, goff):
        assert isinstance(pos, array) and pos.typecode == 'q' and len(pos) == 6
        seen.append(pos)
        return api.entryfunc_abspos(buf, pos, goff)
    got = [list(p) for p in api.readfastq_iter(io.BytesIO(data), 600, entryfunc=spy)]
    assert got == table.tolist()
    assert len({id(p) for p in seen}) == 4  # no aliasing between records (upstream reuses one posbuffer)
    # the module's own entryfunc_abspos is applied to a whole chunk at once: same values, same types, no aliasing
    direct = list(api.readfastq_iter(io.BytesIO(data), 600, entryfunc=api.entryfunc_abspos))
    assert [list(p) for p in direct] == table.tolist() and len({id(p) for p in direct}) == 4
    assert all(isinstance(p, array) and p.typecode == 'q' and len(p) == 6 for p in direct)
    ents = list(api.readfastq_iter(io.BytesIO(data), 600, entryfunc=api.entryfunc))
    assert [e[0] for e in ents] == [data[r[0] + 1:r[1]] for r in table]
    assert [e[1] for e in ents] == [data[r[2]:r[3]] for r in table] and [e[2] for e in ents] == [data[r[4]:r[5]] for r in table]


def test_scan_geometry_follows_the_line_density():
    """device.geometry_for_density: the geometry for the next chunk of a stream from what the last parse saw (host
    logic, no device call): 32 KiB per iteration below 64 bytes per line, the caller's own choice untouched."""
    from fastqandfurious_b200 import device
    gib = 1 << 30
    assert device.geometry_for_density(0, 22_400_000, gib) == 2   # reads wrapped at 60 columns
    assert device.geometry_for_density(0, 19_400_000, gib) == 2   # FASTA-like
    assert device.geometry_for_density(0, 12_744_708, gib) == 0   # 150 bp, four lines per record
    assert device.geometry_for_density(0, 211_520, gib) == 0      # 10 kb reads
    assert device.geometry_for_density(0, 1000, 4096) == 0        # too small to tell
    assert device.geometry_for_density(1, 22_400_000, gib) == 1 and device.geometry_for_density(5, 0, gib) == 5
