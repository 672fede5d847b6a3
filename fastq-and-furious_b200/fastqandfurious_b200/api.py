"""The reference's plugin API for the FASTQ hot path, backed by the B200 kernels.

Mirrors ``fastqandfurious`` (src/fastqandfurious.py) and ``fastqandfurious._fastqandfurious``
(src/_fastqandfurious.c): same names, argument meaning, return values and error messages for
``readfastq_iter`` / ``entrypos`` / ``entryfunc*`` / ``arrayadd_b`` / ``arrayadd_q`` and the status
constants.  The work is done on the GPU in batches: one device call walks the whole entrypos chain of
a chunk; Python only replays the resulting offset table through the caller's ``entryfunc``.
"""
import importlib
import io
import os
import sys
import typing
from array import array
from collections import namedtuple

import numpy as np
import torch

from . import _lib, device
from ._lib import (COMPLETE, INVALID, MISSING_QUAL_BEGIN, MISSING_QUAL_END, MISSING_QUALHEADER_END,  # noqa: F401
                   MISSING_SEQ_BEG, MISSING_SEQ_END, MISSING_SEQHEADER_BEGIN, MISSING_SEQHEADER_END, POS_HEAD_BEG,
                   POS_HEAD_END, POS_QUAL_BEG, POS_QUAL_END, POS_SEQ_BEG, POS_SEQ_END)

Entry = namedtuple('Entry', 'header sequence quality')  # src/fastqandfurious.py:16
DEFAULT_DEVICE_CHUNK = 1 << 26


def read(fh: typing.BinaryIO, fbufsize: int) -> typing.Tuple[bytes, bool]:
    """src/fastqandfurious.py:30-36: (blob, eof) with eof iff the read came back short."""
    blob = fh.read(fbufsize)
    return (blob, len(blob) < fbufsize)


def entryfunc(buf, pos, globaloffset):
    """(header, sequence, quality) slices, src/fastqandfurious.py:161-171."""
    return (buf[(pos[0] + 1):pos[1]], buf[pos[2]:pos[3]], buf[pos[4]:pos[5]])


def entryfunc_namedtuple(buf, pos, globaloffset):
    """Entry(header, sequence, quality), src/fastqandfurious.py:146-158."""
    return Entry(buf[(pos[0] + 1):pos[1]], buf[pos[2]:pos[3]], buf[pos[4]:pos[5]])


def entryfunc_abspos(buf, pos, globaloffset):
    """Absolute stream positions, in place, same object returned (src/fastqandfurious.py:186-195)."""
    for i in (0, 1, 2, 3, 4, 5):
        pos[i] += globaloffset
    return pos


# ---------------------------------------------------------------------------------------------------
# host <-> device staging
# ---------------------------------------------------------------------------------------------------
class _Stager:
    """Pinned host staging buffer + device buffer, grown on demand."""

    def __init__(self, dev):
        self.dev = dev
        self.pinned = None
        self.dbuf = None

    def upload(self, blob):
        n = len(blob)
        if self.pinned is None or self.pinned.numel() < n:
            size = max(n, 1 << 16)
            self.pinned = torch.empty(size, dtype=torch.uint8).pin_memory()
            self.dbuf = torch.empty(size, dtype=torch.uint8, device=self.dev)
        if n:
            self.pinned[:n].numpy()[:] = np.frombuffer(blob, dtype=np.uint8)
            self.dbuf[:n].copy_(self.pinned[:n], non_blocking=True)
        return self.dbuf[:n]


def _device(dev):
    if dev is None:
        if not torch.cuda.is_available():
            raise RuntimeError('fastqandfurious_b200 needs a CUDA device (there is no CPU fallback)')
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(dev)


# ---------------------------------------------------------------------------------------------------
# entrypos: the per-record plugin slot, answered from a chain computed on the device
# ---------------------------------------------------------------------------------------------------
def _byteview(buf):
    mv = memoryview(buf)
    if mv.ndim != 1 or mv.itemsize != 1:
        mv = mv.cast('B')
    return mv


class _ChainCache:
    """Shared machinery of DeviceEntryPos / DeviceEntryPosFasta: per-record calls answered from a chain of calls
    walked on the GPU over a WINDOW of the buffer.

    The reference's functions are stateless (src/_fastqandfurious.c:32,150-151 borrow the buffers for one call);
    the cache must never change an answer:
      * a call is answered from the cache only if it is made on the same buffer object with the same length at an
        offset that is ON the cached chain, and -- for mutable buffers (bytearray, memoryview, array, numpy) -- if
        the bytes the reference would have read for that call, [offset, end of the record + 2) (to the end of
        the buffer for a call that is not COMPLETE), still equal the snapshot that was parsed (the pinned staging
        copy).  Anything else is parsed again, so an in-place rewrite can never return a stale position;
      * a window that does not reach the end of the buffer only yields its COMPLETE calls (positions and status of
        a COMPLETE call do not depend on what follows pos5 + 2; every other status may).  A first call that is not
        COMPLETE inside the window is parsed again with the window doubled until it is, or the buffer ends.
    Windows start at 1 MiB and grow fourfold (up to `max_window`) while the calls keep following the chain, so a
    sequential consumer is served in large batches and a random-access caller pays for 1 MiB per call, not for
    the rest of the buffer.  Calls are serialised by a lock; the buffer reference and the chain are dropped as soon
    as the chain's last answer has been handed out."""

    MIN_WINDOW = 1 << 20
    NEXT_COL = 5    # column of the row whose value (+ NEXT_ADD) is the offset of the next call
    NEXT_ADD = -1
    NPOS = 6

    def __init__(self, device=None, max_window=DEFAULT_DEVICE_CHUNK, min_window=None):
        import threading
        self._dev = device
        self._lock = threading.Lock()
        self._stager = None
        if min_window is not None:
            self.MIN_WINDOW = max(1, int(min_window))
        self._max_window = max(int(max_window), self.MIN_WINDOW)
        self._window = self.MIN_WINDOW
        self._drop()

    def _drop(self):
        self._id = None      # (id(buf), len) of the cached buffer
        self._buf = None     # keeps id(buf) from being reused while the chain is cached
        self._rows = None
        self._next = {}
        self._tail = None    # (offset, status, posbuffer) of the call that ends the chain, or None (window ended first)
        self._lo = self._hi = 0   # window [lo, hi) of the buffer that was parsed
        self._cont = None
        self._mutable = False

    def reset(self):
        """Forget the cached chain."""
        with self._lock:
            self._drop()
            self._window = self.MIN_WINDOW

    def _device_chain(self, d, off):
        raise NotImplementedError

    def _parse(self, buf, mv, offset):
        dev = _device(self._dev)
        if self._stager is None:
            self._stager = _Stager(dev)
        n = len(mv)
        off = self._clamp(offset, n)
        window = self._window
        while True:
            hi = min(n, off + window)
            with torch.cuda.device(dev):
                d = self._stager.upload(mv[off:hi])
                rows, status, tail_pos = self._device_chain(d, off)
            if len(rows) or hi == n or status == INVALID or window >= (1 << 40):
                break
            window *= 2  # the first call is not COMPLETE inside the window: it may be with more bytes
        self._rows = rows
        self._next = {}
        prev = offset
        for k in range(len(rows)):
            self._next[prev] = k
            prev = int(rows[k, self.NEXT_COL]) + self.NEXT_ADD
        # INVALID needs every position of the call inside the window: it does not depend on what follows either
        self._tail = (prev, status, tail_pos) if (hi == n or status == INVALID) else None
        self._cont = prev    # offset of the call that follows the last cached row
        self._id = (id(buf), n)
        self._buf = buf
        self._lo, self._hi = off, hi
        self._mutable = not isinstance(buf, bytes)

    @staticmethod
    def _clamp(offset, n):
        return min(max(int(offset), 0), n)

    def _unchanged(self, mv, lo, hi):
        """Mutable buffers: bytes [lo, hi) still equal the snapshot that was parsed."""
        if not self._mutable:
            return True
        lo, hi = max(lo, self._lo), min(hi, self._hi)
        if hi <= lo:
            return True
        snap = self._stager.pinned[lo - self._lo:hi - self._lo].numpy()
        return bool(np.array_equal(np.frombuffer(mv[lo:hi], dtype=np.uint8), snap))

    def _answer(self, buf, offset):
        """(row, COMPLETE, None) or (None, status, posbuffer) of the call the reference would make at `offset`."""
        mv = _byteview(buf)
        n = len(mv)
        for attempt in (0, 1):
            if self._id == (id(buf), n) and self._buf is buf:
                k = self._next.get(offset)
                if k is not None:
                    row = self._rows[k]
                    if self._unchanged(mv, self._clamp(offset, n), int(row[self.NEXT_COL]) + 2):
                        return row, COMPLETE, None
                elif self._tail is not None and offset == self._tail[0]:
                    if self._unchanged(mv, self._clamp(offset, n), n):
                        _, status, pos = self._tail
                        self._drop()  # the chain is used up: release the buffer
                        return None, status, pos
                elif self._tail is None and offset == self._cont:
                    self._window = min(self._window * 4, self._max_window)  # sequential consumer: larger batches
                else:
                    self._window = self.MIN_WINDOW  # off the chain: random access
            elif attempt == 0 and self._id is not None and self._id[0] != id(buf):
                self._window = max(self.MIN_WINDOW, min(self._window, n))  # another buffer: keep the batch size
            if attempt:
                break
            self._parse(buf, mv, offset)
        raise AssertionError('entrypos chain cache: no answer after a fresh parse')  # a fresh parse holds `offset`


class DeviceEntryPos(_ChainCache):
    """Callable with the contract of ``_fastqandfurious.entrypos(buf, offset, posbuffer) -> status``
    (src/_fastqandfurious.c:25-153): positions relative to ``buf``, posbuffer reset to -1 first, never raises for
    data reasons.  A call that cannot be answered from the cached chain walks the chain from ``offset`` on the GPU
    (over a window of the buffer, see _ChainCache); the calls the reference's loop makes next (offset = pos5 - 1 of
    the previous record, src/fastqandfurious.py:254) are answered from that table.  Answers never depend on the
    cache: mutable buffers are re-validated byte for byte over the span the C function would have read."""

    def _device_chain(self, d, off):
        res = device.parse_buffer(d, sentinel=False, goff=off)
        rows = res.table.cpu().numpy()
        return rows, res.tail_status, [p + off if p >= 0 else -1 for p in res.tail_pos]

    def __call__(self, buf, offset, posbuffer):
        if getattr(posbuffer, 'itemsize', 8) != 8:
            raise ValueError('The buffer must be of format type q.')  # src/_fastqandfurious.c:38-43
        if len(posbuffer) < 6:
            raise ValueError('posbuffer must hold 6 positions')
        with self._lock:
            row, status, pos = self._answer(buf, offset)
        if row is not None:
            for i in range(6):
                posbuffer[i] = int(row[i])
            return COMPLETE
        for i in range(6):
            posbuffer[i] = pos[i]
        return status


entrypos = DeviceEntryPos()


# ---------------------------------------------------------------------------------------------------
# readfastq_iter
# ---------------------------------------------------------------------------------------------------
def _chunks(fh, fbufsize, device_chunk, dev, decode_quality=False, stats=None, base=0):
    """Generator over (blob, rows, qual, goff, final) for successive blobs of the stream, applying the
    refill / end-of-stream rules of src/fastqandfurious.py:256-279 once per blob instead of once per
    record.  rows: int64 ndarray [n,6] relative to blob (the last one already patched at EOF)."""
    stager = _Stager(dev)
    goff = base - 1  # src/fastqandfurious.py:242 sets -1 whatever the argument says; here `globaloffset` is the base
    carry = b'\n'  # :245
    # the first read asks for `fbufsize` bytes like the reference (a pipe or socket yields its first records as soon
    # as that much has arrived), later reads grow fourfold up to the device chunk
    nmax = max(int(fbufsize), int(device_chunk))
    nread = max(1, min(int(fbufsize), nmax))
    table = None
    while True:
        chunk, eof = read(fh, nread)
        nread = min(nmax, nread * 4)
        blob = carry + chunk if carry else chunk
        with torch.cuda.device(dev):
            d = stager.upload(blob)
            res = device.parse_buffer(d, sentinel=False, goff=0, decode_quality=decode_quality, table=table)
            table = res.table_full
            rows = res.table.cpu().numpy()
            qual = res.qual.cpu().numpy() if decode_quality else None
        if stats is not None:
            stats['h2d_bytes'] = stats.get('h2d_bytes', 0) + len(blob)
            stats['d2h_bytes'] = stats.get('d2h_bytes', 0) + rows.nbytes + 128 + (len(blob) if decode_quality else 0)
            stats['device_calls'] = stats.get('device_calls', 0) + 1
        offset = res.resume_offset
        status = res.tail_status
        if eof:
            if status == MISSING_SEQHEADER_BEGIN:
                yield blob, rows, qual, goff
                return
            if status == MISSING_QUAL_END:
                tp = res.tail_pos
                qualend_i = tp[4] + (tp[3] - tp[2])
                if qualend_i >= len(blob):
                    yield blob, rows, qual, goff
                    raise ValueError('Incomplete final quality string at byte')
                last = np.array([tp[0], tp[1], tp[2], tp[3], tp[4], qualend_i], dtype=np.int64)
                if decode_quality:  # the patched span may leave the quality line: decode it explicitly
                    with torch.cuda.device(dev):
                        q = d[tp[4]:qualend_i].clone().view(torch.int8)
                        device.arrayadd_b_(q, -33)
                        qual[tp[4]:qualend_i] = q.cpu().numpy()
                yield blob, np.concatenate([rows, last[None, :]]), qual, goff
                return
            yield blob, rows, qual, goff
            if status == INVALID:
                # the reference spins forever here (no branch of :256-270 matches); raising is the
                # documented decision of this build
                raise ValueError('Entry is invalid at byte %i' % (goff + offset))
            raise ValueError('Incomplete entry at byte %i' % (goff + offset))
        yield blob, rows, qual, goff
        if status == INVALID:
            raise ValueError('Entry is invalid at byte %i' % (goff + offset))
        goff += offset
        carry = blob[offset:]


def _reference_loop(fh, fbufsize, entryfunc, entrypos, base=0):
    """The reference's own per-record loop (src/fastqandfurious.py:241-279) for a caller-supplied
    ``entrypos`` plugin; host logic only."""
    posbuffer = array('q', [-1, ] * 6)
    globaloffset = base - 1
    offset = 0
    buf, eof = read(fh, fbufsize)
    buf = b'\n' + buf
    while True:
        status = entrypos(buf, offset, posbuffer)
        if status == COMPLETE:
            offset = posbuffer[-1] - 1
            yield entryfunc(buf, posbuffer, globaloffset)
        elif eof:
            if status == MISSING_SEQHEADER_BEGIN:
                break
            elif status == MISSING_QUAL_END:
                qualend_i = posbuffer[-2] + (posbuffer[3] - posbuffer[2])
                if qualend_i >= len(buf):
                    raise ValueError('Incomplete final quality string at byte')
                posbuffer[-1] = qualend_i
                yield entryfunc(buf, posbuffer, globaloffset)
                break
            elif status != INVALID:
                raise ValueError('Incomplete entry at byte %i' % (globaloffset + offset))
            else:
                raise ValueError('Entry is invalid at byte %i' % (globaloffset + offset))
        elif status == INVALID:
            raise ValueError('Entry is invalid at byte %i' % (globaloffset + offset))
        else:
            globaloffset += offset
            tmp_buf, eof = read(fh, fbufsize)
            buf = buf[offset:] + tmp_buf
            offset = 0


def readfastq_iter(fh, fbufsize, entryfunc=entryfunc, entrypos=entrypos, globaloffset=0, device=None,
                   device_chunk=DEFAULT_DEVICE_CHUNK, entryfunc_qual=None):
    """Iterate through the entries of a FASTQ stream (drop-in for src/fastqandfurious.py:198-279).

    With the default ``entrypos`` (or any DeviceEntryPos) chunks of max(fbufsize, device_chunk) bytes
    are parsed on the GPU in one call each and the offset table is replayed through ``entryfunc`` --
    results do not depend on the chunk size (neither do the reference's).  Any other ``entrypos``
    callable runs the reference's per-record loop unchanged.

    ``globaloffset``: absolute stream position of the first byte ``fh`` delivers (e.g. ``fh.tell()`` after a seek);
    every ``globaloffset`` handed to ``entryfunc`` -- and so every position ``entryfunc_abspos`` returns and the byte
    quoted in error messages -- is shifted by it.  The reference accepts the argument and then overwrites it with
    -1 (src/fastqandfurious.py:242); the default 0 reproduces exactly that.

    entryfunc_qual(buf, qualbuf, pos, globaloffset): optional; when given it is called instead of
    ``entryfunc`` and ``qualbuf`` is an int8 mirror of ``buf`` whose [pos4:pos5] span holds the
    Phred-33 decoded qualities (the arrayadd_b recipe of src/demo/benchmark.py:161-163, fused)."""
    base = int(globaloffset)
    if not isinstance(entrypos, DeviceEntryPos):
        yield from _reference_loop(fh, fbufsize, entryfunc, entrypos, base)
        return
    dev = _device(device if device is not None else entrypos._dev)
    for blob, rows, qual, goff in _chunks(fh, fbufsize, device_chunk, dev, decode_quality=entryfunc_qual is not None,
                                          base=base):
        flat = array('q')  # one conversion per chunk; every record gets its own 6-item array('q') (a slice)
        if entryfunc is entryfunc_abspos and entryfunc_qual is None:
            # the module's own entryfunc_abspos (pos[i] += globaloffset, the array returned): added for the whole
            # chunk at once, a fresh array('q') per record as below
            flat.frombytes(memoryview(np.ascontiguousarray(rows + goff, dtype=np.int64)).cast('B'))
            for k6 in range(0, 6 * len(rows), 6):
                yield flat[k6:k6 + 6]
            continue
        flat.frombytes(memoryview(np.ascontiguousarray(rows, dtype=np.int64)).cast('B'))
        for k in range(len(rows)):
            pos = flat[6 * k:6 * k + 6]
            if entryfunc_qual is not None:
                yield entryfunc_qual(blob, qual, pos, goff)
            else:
                yield entryfunc(blob, pos, goff)


def _readinto(fh, mv):
    """Fill the writable memoryview `mv` from `fh` (readinto when the object has it, else read + copy).
    Returns the number of bytes obtained (< len(mv) only at the end of the stream)."""
    got, want = 0, len(mv)
    ri = getattr(fh, 'readinto', None)
    while got < want:
        if ri is not None:
            k = ri(mv[got:])
            if not k:
                break
        else:
            data = fh.read(want - got)
            k = len(data)
            if not k:
                break
            mv[got:got + k] = data
        got += k
    return got


class _FastSource:
    """Fills staging buffers from a file object with several threads where the object allows it: a regular file is
    read with os.preadv at explicit offsets, so that several cores copy the page cache at once (4.6 -> 11 GB/s on
    the 16-core GPU box); everything else goes through readinto / read on the calling thread (_readinto).  The
    object's position is kept in step, so the caller may mix its own reads in between.  (An io.BytesIO stays on
    readinto: getbuffer() un-shares the underlying bytes object first, a full copy that costs more than it saves.)"""
    MIN_SLICE = 1 << 22

    def __init__(self, fh, threads=None):
        import stat
        self.fh = fh
        self.kind = 'generic'
        self.pool = None
        self.threads = max(1, threads or int(os.environ.get('FQB_READ_THREADS', 0)) or min(8, os.cpu_count() or 1))
        try:
            if isinstance(fh, (io.BufferedReader, io.FileIO)) and fh.seekable():
                self.fd = fh.fileno()
                if stat.S_ISREG(os.fstat(self.fd).st_mode) and hasattr(os, 'preadv'):
                    self.kind = 'file'
        except (OSError, ValueError, AttributeError):
            self.kind = 'generic'
        if self.kind != 'generic' and self.threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(self.threads)

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None

    def _run(self, job, n):
        """job(lo, hi) over [0, n) in slices, on the pool when there is enough to share."""
        k = min(self.threads, max(1, n // self.MIN_SLICE))
        if self.pool is None or k <= 1:
            job(0, n)
            return
        step = -(-n // k)
        for f in [self.pool.submit(job, lo, min(n, lo + step)) for lo in range(0, n, step)]:
            f.result()

    def readinto(self, mv):
        """Fill the writable byte memoryview `mv`; returns the bytes obtained (< len(mv) only at the end)."""
        if self.kind == 'file':
            fh, fd = self.fh, self.fd
            pos = fh.tell()
            n = max(0, min(len(mv), os.fstat(fd).st_size - pos))
            short = []

            def job(lo, hi):
                at = lo
                while at < hi:
                    k = os.preadv(fd, [mv[at:hi]], pos + at)
                    if k <= 0:  # the file shrank under us
                        short.append(at)
                        return
                    at += k
            self._run(job, n)
            if short:
                n = min(short)
            fh.seek(pos + n)
            return n
        return _readinto(self.fh, mv)


def _table_stream(fh, fbufsize, device_chunk, dev, stats=None, base=0):
    """Generator over int64 [n,6] arrays of ABSOLUTE offsets, chunk by chunk, for readfastq_table: the refill and
    end-of-stream rules of src/fastqandfurious.py:256-279 applied once per chunk, with the host side pipelined --
    a reader thread fills one pinned staging buffer straight from the file object (_FastSource: several threads
    for in-memory streams and regular files, readinto otherwise: no intermediate bytes objects) while the other
    one is copied to the device and parsed; the unfinished tail of a chunk
    (src/fastqandfurious.py:274-279: ``buf = buf[offset:] + tmp``) is carried into the head room in front of the
    next chunk's bytes."""
    import queue
    import threading
    chunk = max(int(fbufsize), int(device_chunk))
    room = max(4096, min(1 << 20, chunk))  # head room for the carried tail (more than that: the slow path below)
    key = (str(dev), room + chunk)
    cached = _stream_bufs.pop(key, None)  # staging buffers are kept between calls (pinning is slow)
    if cached is None:
        with torch.cuda.device(dev):
            cached = ([torch.empty(room + chunk, dtype=torch.uint8).pin_memory() for _ in range(2)],
                      torch.empty(room + chunk, dtype=torch.uint8, device=dev))
    pins, dbuf = cached
    views = [p.numpy() for p in pins]
    free, ready = queue.Queue(), queue.Queue()
    free.put(0)
    free.put(1)

    source = _FastSource(fh)

    def reader():
        try:
            while True:
                i = free.get()
                if i is None:
                    return
                n = source.readinto(memoryview(views[i])[room:room + chunk])
                ready.put((i, n, n < chunk))
                if n < chunk:
                    return
        except BaseException as e:  # handed to the consumer
            ready.put(e)

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    carry = b'\n'  # src/fastqandfurious.py:245
    goff = int(base) - 1  # :242 (base = the caller's globaloffset, 0 upstream)
    table = None
    stager = None
    try:
        while True:
            item = ready.get()
            if isinstance(item, BaseException):
                raise item
            i, n, eof = item
            c = len(carry)
            with torch.cuda.device(dev):
                if c <= room:
                    views[i][room - c:room] = np.frombuffer(carry, dtype=np.uint8)
                    d = dbuf[:c + n]
                    d.copy_(pins[i][room - c:room + n], non_blocking=True)
                    tail_of = lambda off: bytes(views[i][room - c + off:room + n])  # noqa: E731
                else:  # a tail longer than the head room (an entry larger than the chunk): plain concatenation
                    blob = carry + bytes(views[i][room:room + n])
                    if stager is None:
                        stager = _Stager(dev)
                    d = stager.upload(blob)
                    tail_of = lambda off: blob[off:]  # noqa: E731
                blob_len = c + n
                res = device.parse_buffer(d, sentinel=False, goff=0, table=table)
                table = res.table_full
                rows = res.table.cpu().numpy()
            if stats is not None:
                stats['h2d_bytes'] = stats.get('h2d_bytes', 0) + blob_len
                stats['d2h_bytes'] = stats.get('d2h_bytes', 0) + rows.nbytes + 128
                stats['device_calls'] = stats.get('device_calls', 0) + 1
            offset, status = res.resume_offset, res.tail_status
            if eof:
                if status == MISSING_SEQHEADER_BEGIN:
                    yield rows + goff
                    return
                if status == MISSING_QUAL_END:
                    tp = res.tail_pos
                    qualend_i = tp[4] + (tp[3] - tp[2])
                    yield rows + goff
                    if qualend_i >= blob_len:
                        raise ValueError('Incomplete final quality string at byte')
                    yield np.array([[tp[0], tp[1], tp[2], tp[3], tp[4], qualend_i]], dtype=np.int64) + goff
                    return
                yield rows + goff
                if status == INVALID:
                    raise ValueError('Entry is invalid at byte %i' % (goff + offset))
                raise ValueError('Incomplete entry at byte %i' % (goff + offset))
            yield rows + goff
            if status == INVALID:
                raise ValueError('Entry is invalid at byte %i' % (goff + offset))
            carry = tail_of(offset)
            goff += offset
            free.put(i)
    finally:
        free.put(None)
        # normal end: the reader has returned already; after an error it may sit in a blocking read -- do not wait
        # for it (daemon thread; the staging buffers are then not reused)
        th.join(timeout=5 if sys.exc_info()[0] is None else 0.05)
        if not th.is_alive():
            source.close()
        if not th.is_alive() and len(_stream_bufs) < 4:
            _stream_bufs[key] = cached


_stream_bufs = {}


def readfastq_table(fh, fbufsize=2 ** 16, device=None, device_chunk=DEFAULT_DEVICE_CHUNK, stats=None, globaloffset=0):
    """All of ``readfastq_iter(fh, fbufsize, entryfunc=entryfunc_abspos)`` at once: int64 ndarray [n,6]
    of absolute stream offsets (the on-disk index of src/demo/benchmark.py:268-287).  Raises the same
    ValueErrors; rows parsed before the error are attached to the exception as ``.rows``.  The file object is
    read by a background thread into pinned staging buffers while the previous chunk is on the GPU."""
    dev = _device(device)
    parts = []
    try:
        for rows in _table_stream(fh, fbufsize, device_chunk, dev, stats=stats, base=globaloffset):
            parts.append(rows)
    except ValueError as e:
        e.rows = np.concatenate(parts) if parts else np.empty((0, 6), dtype=np.int64)
        raise
    return np.concatenate(parts) if parts else np.empty((0, 6), dtype=np.int64)


# ---------------------------------------------------------------------------------------------------
# arrayadd_b / arrayadd_q on host arrays (device round trip) and on CUDA tensors (in place)
# ---------------------------------------------------------------------------------------------------
def _arrayadd(a, value, kind):
    if isinstance(a, torch.Tensor):  # CUDA tensor: in place, asynchronous; the reference returns None
        (device.arrayadd_b_ if kind == 'b' else device.arrayadd_q_)(a, value)
        return None
    mv = memoryview(a)
    want = 1 if kind == 'b' else 8
    if mv.itemsize != want:
        raise ValueError('The buffer must be of format type %s.' % kind)
    if mv.readonly:
        raise TypeError('a writable buffer is required')
    if mv.nbytes == 0:
        return None
    dt = np.int8 if kind == 'b' else np.int64
    host = np.frombuffer(mv, dtype=dt)
    dev = _device(None)
    with torch.cuda.device(dev):
        t = torch.from_numpy(host.copy()).to(dev)
        (device.arrayadd_b_ if kind == 'b' else device.arrayadd_q_)(t, value)
        host[:] = t.cpu().numpy()
    return None


def arrayadd_b(a, value):
    """a[i] += (int8)value in place with wrap (src/_fastqandfurious.c:161-185)."""
    return _arrayadd(a, value, 'b')


def arrayadd_q(a, value):
    """a[i] += value in place on int64 items (src/_fastqandfurious.c:193-217)."""
    return _arrayadd(a, value, 'q')


# ---------------------------------------------------------------------------------------------------
# FASTA (src/fastqandfurious.py:103-143)
# ---------------------------------------------------------------------------------------------------
class DeviceEntryPosFasta(_ChainCache):
    """Callable with the contract of ``entrypos_fasta(buf, offset, posbuffer) -> status``
    (src/fastqandfurious.py:103-143): positions relative to ``buf``; like the reference, only the entries that
    were found are assigned (posbuffer is not reset).  Same caching rules as DeviceEntryPos (_ChainCache): the
    natural next calls (offset = pos3 of the previous record) are answered from the chain walked on the GPU, and
    an answer never depends on the cache."""
    NEXT_COL = 3
    NEXT_ADD = 0
    NPOS = 4

    @staticmethod
    def _clamp(offset, n):
        off = int(offset)
        if off < 0:  # bytes.find: a negative start counts from the end
            off = max(0, off + n)
        return min(off, n)

    def _device_chain(self, d, off):
        res = device.parse_fasta_buffer(d, sentinel=False, goff=off)
        rows = res.table.cpu().numpy()
        return rows, res.tail_status, [p + off if p >= 0 else -1 for p in res.tail_pos]

    def __call__(self, buf, offset, posbuffer):
        with self._lock:
            row, status, pos = self._answer(buf, offset)
        if row is not None:
            for i in range(4):
                posbuffer[i] = int(row[i])
            return COMPLETE
        for i in range(4):
            if pos[i] >= 0:
                posbuffer[i] = pos[i]
        return status


entrypos_fasta = DeviceEntryPosFasta()


def entryfunc_fasta(buf, pos, globaloffset):
    """(header, sequence) slices -- the function the reference's FASTA test calls (tests.py:101-105) but the
    module never defined; the FASTA counterpart of entryfunc (src/fastqandfurious.py:161-171)."""
    return (buf[(pos[0] + 1):pos[1]], buf[pos[2]:pos[3]])


# ---------------------------------------------------------------------------------------------------
# automagic_open (host I/O, src/fastqandfurious.py:282-334)
# ---------------------------------------------------------------------------------------------------
FORMAT_OPENERS = {
    'gz': ('gzip', 'open', list()),
    'gzip': ('gzip', 'open', list()),
    'bz2': ('bz2', 'open', list()),
    'lzma': ('lzma', 'open', list()),
}


def automagic_open(filename, openers=None):
    """Open a (possibly compressed) FASTQ file by its extension, as src/fastqandfurious.py:290-334 documents it:
    'foo/bar.fq.gz' -> gzip, '.bz2' -> bz2, '.lzma' -> lzma, anything else -> a plain binary file.  `openers`
    overrides the module-level FORMAT_OPENERS mapping extension -> (module name or namespace, function name,
    positional arguments).  Two upstream defects are not reproduced: the `openers` argument is ignored there
    (:326 reads FORMAT_OPENERS), and `importlib.importmodule` (:330) does not exist, so every built-in
    compressed format raises AttributeError upstream.  The returned stream feeds readfastq_iter / HostParser:
    decompression stays on the host."""
    table = FORMAT_OPENERS if openers is None else openers
    maybe_ext = filename.rsplit(os.path.extsep, maxsplit=1)
    ext = None if len(maybe_ext) == 1 else maybe_ext[-1]
    try:
        modulename, funcname, args = table[ext]
    except KeyError:
        modulename, funcname, args = ('io', 'open', ('rb',))
    module = importlib.import_module(modulename) if isinstance(modulename, str) else modulename
    return getattr(module, funcname)(filename, *args)
