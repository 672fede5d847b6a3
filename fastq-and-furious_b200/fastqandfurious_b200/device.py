"""Device-side entry points: torch owns memory and streams, libfqb200.so does the work.

``parse_buffer`` is the batched replacement of the reference's per-record loop
(``readfastq_iter`` body, src/fastqandfurious.py:251-255, with the C ``entrypos``,
src/_fastqandfurious.c:25-153, and ``entryfunc_abspos``, src/fastqandfurious.py:186-195)."""
import ctypes
from collections import namedtuple

import numpy as np
import torch

from . import _lib
from ._lib import INVALID, MISSING_QUAL_END, MISSING_SEQHEADER_BEGIN

ParseResult = namedtuple('ParseResult',
                         'table n tail_status tail_pos resume_offset path n_lines qual first_bad table_full spec')

_ws_cache = {}
launch_count = 0  # kernels enqueued by this module (bench.py reports it)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError('%s must be a CUDA tensor' % name)
    if not t.is_contiguous():
        raise ValueError('%s must be contiguous' % name)


def _workspace(dev, nbytes):
    # one workspace per (device, stream): parses enqueued on different streams may overlap on the GPU
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _ws_cache.pop(key, None)
        ws = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def parse_raw(buf, sentinel, goff, table, qual, qual_add, result, flags, max_lines=0):
    """One fqb_parse call (asynchronous).  All tensors are caller-owned CUDA tensors."""
    global launch_count
    L = _lib.lib()
    n = buf.numel()
    need = L.fqb_workspace_bytes(n, max_lines, int(flags))
    ws = _workspace(buf.device, need)
    cap = table.shape[0] if table is not None else 0
    code = L.fqb_parse(buf.data_ptr() if n else None, n, int(bool(sentinel)), int(goff),
                       table.data_ptr() if cap else None, cap,
                       qual.data_ptr() if qual is not None else None, int(qual_add), result.data_ptr(),
                       ws.data_ptr(), ws.numel(), int(max_lines), int(flags), _stream())
    _lib.check(code, 'fqb_parse')
    general = max_lines > 0 and not (flags & (_lib.FLAG_FAST_ONLY | _lib.FLAG_SPEC_ONLY))
    spec_pass = not (flags & (_lib.FLAG_FAST_ONLY | _lib.FLAG_NO_SPEC)) and (general or bool(flags & _lib.FLAG_SPEC_ONLY))
    fast = not (flags & _lib.FLAG_FORCE_GENERAL)
    # scan + emit (+ decode kernel for unaligned mirrors) | + speculative pass | + 11 general kernels (+ its decode)
    launch_count += (2 + (1 if (qual is not None and fast and n) else 0) + (1 if spec_pass else 0) +
                     ((11 + (1 if qual is not None else 0)) if general else 0))
    return ws


def geometry_for_density(cfg, n_lines, nbytes):
    """Scan geometry for the NEXT parse of the same stream, from what the last one saw: with lines shorter than
    64 bytes on average (wrapped multi-line records, very short reads) the 32 KiB-per-iteration geometry
    (FQB_FLAG_CFG(2)) scans 4 % faster than the default, which is the faster one at 150 bp and beyond.
    A caller-chosen geometry (cfg != 0) is kept."""
    if cfg == 0 and nbytes > (1 << 20) and int(n_lines) * 64 > int(nbytes):
        return 2
    return cfg


def read_result(result):
    """Device result header -> FqbResult (synchronises the current stream)."""
    host = result.cpu().numpy().tobytes()
    return _lib.FqbResult.from_buffer_copy(host)


def _run(buf, sentinel, goff, table, qual, qual_add, cfg, force_general, spec=True):
    """Fast path, then (only if it declined) the general path.  Returns the FqbResult of the pass that
    produced the answer; ERR_CAPACITY is left to the caller (n_records is exact in that case)."""
    result = torch.empty(16, dtype=torch.int64, device=buf.device)
    flags = _lib.FLAG_CFG(cfg) | (0 if spec else _lib.FLAG_NO_SPEC) | (_lib.FLAG_SPEC_V1 if spec == 'v1' else 0)
    res = None
    # first call: the 4-line fast path and, where it declines, the general path's speculative single pass (both on
    # the device, back to back, no line table); only what neither can answer needs the exact resolution below
    first = (_lib.FLAG_SPEC_ONLY if spec else _lib.FLAG_FAST_ONLY) | (_lib.FLAG_FORCE_GENERAL if force_general else 0)
    if spec or not force_general:
        parse_raw(buf, sentinel, goff, table, qual, qual_add, result, flags | first)
        res = read_result(result)
        if res.error == _lib.ERR_DENSE:  # very short lines: lists sized for one newline per byte
            flags |= _lib.FLAG_DENSE
            parse_raw(buf, sentinel, goff, table, qual, qual_add, result, flags | first)
            res = read_result(result)
        if not res.need_general:
            return res
    max_lines = (res.n_lines + 64) if res is not None else buf.numel() // 32 + 64
    for _ in range(4):
        parse_raw(buf, sentinel, goff, table, qual, qual_add, result, flags | _lib.FLAG_FORCE_GENERAL | _lib.FLAG_NO_SPEC,
                  max_lines=max_lines)
        res = read_result(result)
        if res.error == _lib.ERR_DENSE:
            flags |= _lib.FLAG_DENSE
            continue
        if res.error != _lib.ERR_WORKSPACE:
            break
        max_lines = res.n_lines + 64
    return res


def parse_buffer(buf, sentinel=True, goff=-1, decode_quality=False, qual_add=-33, cap=None, cfg=0,
                 force_general=False, table=None, qual=None, spec=True):
    """Walk the reference's entrypos chain over a device-resident byte buffer.

    buf: uint8 CUDA tensor holding raw FASTQ bytes.  With ``sentinel`` the blob the reference would
    see is ``b'\\n' + buf`` (what readfastq_iter builds for the first chunk, src/fastqandfurious.py:245)
    and ``goff=-1`` turns blob positions into offsets in ``buf`` (entryfunc_abspos semantics).

    Returns ParseResult: ``table`` int64[n,6] = [pos0..pos5]+goff of the COMPLETE records, the status /
    posbuffer / offset of the first call that was not COMPLETE, and (decode_quality) ``qual``, an int8
    mirror of ``buf`` where qual[pos4:pos5] of every record holds byte + qual_add (arrayadd_b recipe,
    src/demo/benchmark.py:161-163); other bytes of ``qual`` are unspecified.

    ``force_general`` skips the 4-line fast path; ``spec=False`` also skips the general path's speculative single
    pass, ``spec='v1'`` runs that pass as one CTA per chunk (csrc/fq_gspec.cuh) instead of one warp per chunk
    (csrc/fq_gspec2.cuh) (tests: every path must give the same table).  ``ParseResult.spec`` tells whether that pass
    answered."""
    _require_cuda(buf, 'buf')
    if buf.dtype != torch.uint8:
        raise TypeError('buf must be uint8')
    dev = buf.device
    n = buf.numel()
    with torch.cuda.device(dev):
        if cap is None:
            cap = n // 96 + 64
        if table is None or table.shape[0] < cap:
            table = torch.empty((cap, 6), dtype=torch.int64, device=dev)
        if decode_quality and qual is None:
            qual = torch.empty(n, dtype=torch.int8, device=dev)
        if not decode_quality:
            qual = None
        if qual is not None:
            _require_cuda(qual, 'qual')
            if qual.dtype != torch.int8 or qual.numel() < n or qual.device != dev:
                raise ValueError('qual must be an int8 CUDA tensor on the buffer\'s device with at least len(buf) elements')
        if table is not None and (table.dtype != torch.int64 or table.dim() != 2 or table.shape[1] != 6 or
                                  table.device != dev or not table.is_contiguous()):
            raise ValueError('table must be a contiguous int64 [cap,6] CUDA tensor on the buffer\'s device')
        for _ in range(3):
            res = _run(buf, sentinel, goff, table, qual, qual_add, cfg, force_general, spec)
            if res.error != _lib.ERR_CAPACITY:
                break
            table = torch.empty((res.n_records + 64, 6), dtype=torch.int64, device=dev)
        if res.error != _lib.ERR_OK:
            raise RuntimeError('fqb_parse: error %d (n_lines=%d)' % (res.error, res.n_lines))
        nrec = res.n_records
        return ParseResult(table[:nrec], nrec, res.tail_status, list(res.tail_pos), res.resume_offset, res.path,
                           res.n_lines, qual, res.first_bad, table, bool(res.path == _lib.PATH_GENERAL and res.reserved[1] == 1))


FastaResult = namedtuple('FastaResult', 'table n tail_status tail_pos resume_offset n_lines')


def parse_fasta_buffer(buf, sentinel=True, goff=-1, cap=None, cfg=0, max_lines=None):
    """Walk the chain of the reference's ``entrypos_fasta`` calls (src/fastqandfurious.py:103-143) over a
    device-resident byte buffer, each call starting at pos3 of the previous record.

    Returns FastaResult: ``table`` int64[n,4] = [pos0 '>', pos1 header end, pos2 sequence start, pos3 sequence
    end] + goff of the COMPLETE calls, and the status / positions (-1 = not assigned by the reference) / offset of
    the first call that is not COMPLETE.  With ``sentinel`` the blob is ``b'\\n' + buf`` (a '>' in the first byte
    starts a record) and ``goff=-1`` turns blob positions into offsets in ``buf``."""
    global launch_count
    _require_cuda(buf, 'buf')
    if buf.dtype != torch.uint8:
        raise TypeError('buf must be uint8')
    L = _lib.lib()
    dev = buf.device
    n = buf.numel()
    flags = _lib.FLAG_CFG(cfg)
    with torch.cuda.device(dev):
        if max_lines is None:
            max_lines = n // 40 + 1024
        if cap is None:
            cap = n // 64 + 64
        result = torch.empty(16, dtype=torch.int64, device=dev)
        for _ in range(6):
            table = torch.empty((cap, 4), dtype=torch.int64, device=dev)
            ws = _workspace(dev, L.fqb_fasta_workspace_bytes(n, max_lines, flags))
            _lib.check(L.fqb_parse_fasta(buf.data_ptr() if n else None, n, int(bool(sentinel)), int(goff),
                                         table.data_ptr(), cap, result.data_ptr(), ws.data_ptr(), ws.numel(),
                                         int(max_lines), int(flags), _stream()), 'fqb_parse_fasta')
            launch_count += 10  # scan, flags, two running-maximum levels, fix-up, three of the prefix sum, rows, result
            res = read_result(result)
            if res.error == _lib.ERR_WORKSPACE:
                max_lines = res.n_lines + 64
            elif res.error == _lib.ERR_DENSE:
                flags |= _lib.FLAG_DENSE
            elif res.error == _lib.ERR_CAPACITY:
                cap = res.n_records + 64
            else:
                break
        if res.error != _lib.ERR_OK:
            raise RuntimeError('fqb_parse_fasta: error %d (n_lines=%d)' % (res.error, res.n_lines))
        return FastaResult(table[:res.n_records], res.n_records, res.tail_status, list(res.tail_pos)[:4],
                           res.resume_offset, res.n_lines)


class HostParser:
    """Whole-stream parse of a HOST buffer (the end-to-end path): chunked, pipelined host->device copies
    on a copy stream, one fqb_parse per chunk as its bytes land, offset rows streamed back to pinned
    host memory on a third stream.  Applies readfastq_iter's refill and end-of-stream rules
    (src/fastqandfurious.py:256-279) at chunk granularity; rows are absolute stream offsets
    (entryfunc_abspos).  Buffers are kept between calls."""

    def __init__(self, device='cuda', chunk_bytes=1 << 27, cfg=0):
        self.dev = torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device('cuda', torch.cuda.current_device())
        self.chunk = int(chunk_bytes)
        self.cfg = cfg
        self.dbuf = None
        self.dtable = None
        self.htable = None
        with torch.cuda.device(self.dev):
            self.copy_stream = torch.cuda.Stream()
            self.d2h_stream = torch.cuda.Stream()
        self.stats = {}

    def _ensure(self, nbytes, rows):
        if self.dbuf is None or self.dbuf.numel() < nbytes:
            self.dbuf = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        if self.dtable is None or self.dtable.shape[0] < rows:
            old_d, old_h = self.dtable, self.htable
            self.dtable = torch.empty((rows, 6), dtype=torch.int64, device=self.dev)
            self.htable = torch.empty((rows, 6), dtype=torch.int64).pin_memory()
            return old_d, old_h
        return None, None

    def parse(self, host):
        """host: 1-D uint8 CPU tensor (pinned memory gives full PCIe speed).  Returns int64 ndarray
        [n,6] (a view of pinned memory owned by this object, valid until the next call)."""
        if not isinstance(host, torch.Tensor) or host.is_cuda or host.dtype != torch.uint8 or host.dim() != 1:
            raise TypeError('host must be a 1-D uint8 CPU tensor')
        total = host.numel()
        with torch.cuda.device(self.dev):
            self._ensure(max(total, 16), total // 96 + 64)
            cur = torch.cuda.current_stream()
            self.copy_stream.wait_stream(cur)
            bounds = list(range(0, total, self.chunk)) + [total]
            if total == 0:
                bounds = [0, 0]
            events = []
            with torch.cuda.stream(self.copy_stream):
                for c in range(len(bounds) - 1):
                    lo, hi = bounds[c], bounds[c + 1]
                    if hi > lo:
                        self.dbuf[lo:hi].copy_(host[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
                    events.append(ev)
            start, sentinel, ntot = 0, 1, 0
            status, tail, resume_abs = MISSING_SEQHEADER_BEGIN, [-1] * 6, -1
            ncalls = 0
            run_cfg = self.cfg  # the chunks behind the first one run the scan geometry its line density asks for
            for c in range(len(bounds) - 1):
                cur.wait_event(events[c])
                end = bounds[c + 1]
                goff = start - sentinel
                while True:
                    res = _run(self.dbuf[start:end], sentinel, goff, self.dtable[ntot:], None, 0, run_cfg, False)
                    ncalls += 1
                    if res.error == _lib.ERR_CAPACITY:
                        old_d, _ = self._ensure(0, 2 * (ntot + res.n_records) + 1024)
                        self.d2h_stream.synchronize()
                        self.dtable[:ntot].copy_(old_d[:ntot])
                        self.htable[:ntot].copy_(old_d[:ntot])
                        continue
                    if res.error != _lib.ERR_OK:
                        raise RuntimeError('fqb_parse: error %d' % res.error)
                    break
                n = res.n_records
                run_cfg = geometry_for_density(self.cfg, res.n_lines, end - start)
                if n:
                    self.d2h_stream.wait_stream(cur)
                    with torch.cuda.stream(self.d2h_stream):
                        self.htable[ntot:ntot + n].copy_(self.dtable[ntot:ntot + n], non_blocking=True)
                    ntot += n
                status = res.tail_status
                tail = [p + goff if p >= 0 else -1 for p in res.tail_pos]
                resume_abs = goff + res.resume_offset  # what the reference prints in its error messages
                if status == INVALID and c < len(bounds) - 2:
                    break
                if res.resume_offset > 0:
                    start = start + res.resume_offset - sentinel
                    sentinel = 0
            self.d2h_stream.synchronize()
            self.stats = {'h2d_bytes': total, 'd2h_bytes': ntot * 48 + 128 * ncalls, 'device_calls': ncalls}
            rows = self.htable[:ntot].numpy()
            # end-of-stream rules, src/fastqandfurious.py:256-273
            if status == MISSING_SEQHEADER_BEGIN:
                return rows
            if status == MISSING_QUAL_END and end == total:
                qualend = tail[4] + (tail[3] - tail[2])
                if qualend >= total:  # absolute: blob length - 1 == total
                    err = ValueError('Incomplete final quality string at byte')
                    err.rows = rows
                    raise err
                self.htable[ntot] = torch.tensor(tail[:5] + [qualend], dtype=torch.int64)
                return self.htable[:ntot + 1].numpy()
            err = ValueError(('Entry is invalid at byte %i' if status == INVALID else 'Incomplete entry at byte %i')
                             % resume_abs)
            err.rows = rows
            raise err


def arrayadd_b_(t, value):
    """In-place int8 add with wrap on a CUDA tensor (arrayadd_b, src/_fastqandfurious.c:161-185)."""
    global launch_count
    _require_cuda(t, 'tensor')
    if t.dtype not in (torch.int8, torch.uint8):
        raise ValueError('The buffer must be of format type b.')
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().fqb_arrayadd_b(t.data_ptr() if t.numel() else None, t.numel(), _short_to_byte(value),
                                             _stream()), 'fqb_arrayadd_b')
    launch_count += 1
    return t


def arrayadd_q_(t, value):
    """In-place int64 add on a CUDA tensor (arrayadd_q, src/_fastqandfurious.c:193-217)."""
    global launch_count
    _require_cuda(t, 'tensor')
    if t.dtype != torch.int64:
        raise ValueError('The buffer must be of format type q.')
    value = int(value)
    if not -2 ** 63 <= value < 2 ** 63:
        raise OverflowError('value does not fit a signed 64-bit integer')
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().fqb_arrayadd_q(t.data_ptr() if t.numel() else None, t.numel(), value, _stream()),
                   'fqb_arrayadd_q')
    launch_count += 1
    return t


def _short_to_byte(value):
    """The reference parses `value` as a C short ("h", src/_fastqandfurious.c:167) and keeps its low byte."""
    value = int(value)
    if not -2 ** 15 <= value < 2 ** 15:
        raise OverflowError('signed short integer is out of range')
    return value & 0xff


def synth_fixed(n_records, header_len=32, read_len=150, seed=0xB2000002, device='cuda', first_byte=0, n_bytes=None):
    """Fixed-geometry synthetic FASTQ generated on the device (bench / full-size parity tests): bytes
    [first_byte, first_byte + n_bytes) of an unbounded stream of records (default: n_records whole ones)."""
    global launch_count
    rec = header_len + 1 + read_len + 1 + 2 + read_len + 1
    if n_bytes is None:
        n_bytes = n_records * rec
    dev = torch.device(device)
    with torch.cuda.device(dev):
        buf = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib().fqb_synth_fixed(buf.data_ptr() if n_bytes else None, n_bytes, first_byte, header_len,
                                              read_len, seed, _stream()), 'fqb_synth_fixed')
    launch_count += 1
    return buf


def kernel_info(cfg=0):
    a, b, c, d = (ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32())
    _lib.check(_lib.lib().fqb_kernel_info(cfg, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)),
               'fqb_kernel_info')
    return {'tile_bytes': a.value, 'threads': b.value, 'stages': c.value, 'ctas_per_sm': d.value}


def bind_host_to_gpu(dev=None):
    """Pin the calling process to the CPUs that are local to the GPU (its PCIe root's NUMA node), so that pinned
    staging buffers allocated afterwards are first-touched next to the device.  With one process per GPU this
    keeps 8 concurrent host->device streams from crossing the socket interconnect.  Returns the CPU set used,
    or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        d = torch.device('cuda', torch.cuda.current_device()) if dev is None else torch.device(dev)
        props = torch.cuda.get_device_properties(d)
        bdf = '%04x:%02x:%02x.0' % (getattr(props, 'pci_domain_id', 0), props.pci_bus_id, props.pci_device_id)
        with open('/sys/bus/pci/devices/%s/local_cpulist' % bdf) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
