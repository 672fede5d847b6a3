"""Synthetic FASTQ streams with variable record geometry, generated on the device together with their TRUE offset
table (BASELINE.json configs[2..4]): 'illumina' (150 bp, variable-width headers), 'ont' (long reads,
10 kb mean) and 'multiline' (reads wrapped at 60 columns, long '+' lines).  bench.py and the full-size parity tests
use it; it is not part of the reference's API.  The byte-level definition lives in csrc/fq_synth.cuh, the numpy twin in
tests/fqgen.py (synth_records_np)."""
import ctypes
import math

import numpy as np
import torch

from . import _lib, device

KINDS = {'illumina': _lib.SYNTH_ILLUMINA, 'ont': _lib.SYNTH_ONT, 'multiline': _lib.SYNTH_MULTILINE}
SEEDS = {'illumina': 0xB2000003, 'ont': 0xB2000004, 'multiline': 0xB2000005}
QT_BITS = 12

_qt_cache = {}


def ont_qtable(theta=5000.0, lo=200, hi=500000):
    """Quantile table of the ONT-like read lengths: entry i = clip(F^-1(i / 4096), lo, hi) for the Gamma(k=2, theta)
    distribution (mean 2 theta = 10 kb), F(x) = 1 - (1 + x/theta) exp(-x/theta), inverted by bisection.  4097 int32
    entries; the generators interpolate linearly inside a bin with integer arithmetic, so the device and the numpy
    twin agree bit for bit whatever the floating-point details of this table are."""
    key = (theta, lo, hi)
    if key not in _qt_cache:
        n = 1 << QT_BITS
        out = np.empty(n + 1, dtype=np.int32)
        for i in range(n + 1):
            p = i / n
            if p >= 1.0:
                x = float('inf')
            else:
                a, b = 0.0, 64.0
                for _ in range(80):
                    m = 0.5 * (a + b)
                    if 1.0 - (1.0 + m) * math.exp(-m) < p:
                        a = m
                    else:
                        b = m
                x = 0.5 * (a + b) * theta
            out[i] = int(min(max(x, lo), hi))
        _qt_cache[key] = out
    return _qt_cache[key]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class SynthStream:
    """One synthetic stream: records 0 .. n-1 of `kind`; `off` (device int64 [n+1]) holds the stream offset of every
    record (= its pos0) and the total length."""

    def __init__(self, kind, n_records, seed=None, dev='cuda', qtable=None):
        self.kind_name = kind
        self.kind = KINDS[kind]
        self.seed = SEEDS[kind] if seed is None else int(seed)
        self.dev = torch.device(dev)
        if self.dev.index is None:
            self.dev = torch.device('cuda', torch.cuda.current_device())
        self.n = int(n_records)
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            if kind == 'ont':  # read-length quantiles: 4097 ascending int32 (default: the 10 kb-mean Gamma of cfg 4)
                qt = np.ascontiguousarray(ont_qtable() if qtable is None else qtable, dtype=np.int32)
                if qt.shape != ((1 << QT_BITS) + 1,) or (np.diff(qt) < 0).any() or qt[0] < 1:
                    raise ValueError('qtable: %d ascending positive int32 entries' % ((1 << QT_BITS) + 1))
                self.qtable = torch.from_numpy(qt.copy()).to(self.dev)
            else:
                self.qtable = None
            self.off = torch.zeros(self.n + 1, dtype=torch.int64, device=self.dev)
            if self.n:
                _lib.check(L.fqb_synth_meta(self.kind, self.seed, 0, self.n, self._qt(), None, self.off.data_ptr(),
                                            _stream()), 'fqb_synth_meta')
                ws = torch.empty(L.fqb_scan_workspace_bytes(self.n) + 256, dtype=torch.uint8, device=self.dev)
                _lib.check(L.fqb_exclusive_scan(self.off.data_ptr(), self.n, self.off.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _stream()), 'fqb_exclusive_scan')
                device.launch_count += 4
            self.total = int(self.off[-1].item())

    def _qt(self):
        return self.qtable.data_ptr() if self.qtable is not None else None

    @classmethod
    def for_bytes(cls, kind, target_bytes, seed=None, dev='cuda', qtable=None):
        """The longest stream of whole records that fits `target_bytes`."""
        probe = cls(kind, 1 << 16, seed, dev, qtable)
        mean = probe.total / probe.n
        guess = int(target_bytes / mean * 1.02) + 1024
        s = cls(kind, guess, seed, dev, qtable)
        while s.total < target_bytes:  # the estimate fell short (very skewed lengths)
            guess = int(guess * 1.1) + 1024
            s = cls(kind, guess, seed, dev, qtable)
        n = int(torch.searchsorted(s.off, torch.tensor([int(target_bytes)], device=s.dev), right=True).item()) - 1
        s.n = n
        s.off = s.off[:n + 1].clone()
        s.total = int(s.off[-1].item())
        return s

    def fill(self, first_byte=0, n_bytes=None, out=None):
        """Bytes [first_byte, first_byte + n_bytes) of the stream as a uint8 CUDA tensor."""
        if n_bytes is None:
            n_bytes = self.total - first_byte
        if first_byte < 0 or n_bytes < 0 or first_byte + n_bytes > self.total:
            raise ValueError('window outside the stream')
        with torch.cuda.device(self.dev):
            if out is None:
                out = torch.empty(n_bytes, dtype=torch.uint8, device=self.dev)
            if out.numel() < n_bytes or out.dtype != torch.uint8 or not out.is_contiguous():
                raise ValueError('out must be a contiguous uint8 tensor of at least n_bytes elements')
            _lib.check(_lib.lib().fqb_synth_fill(self.kind, self.seed, 0, self.n, self.off.data_ptr(), self._qt(),
                                                 out.data_ptr() if n_bytes else None, first_byte, n_bytes, _stream()),
                       'fqb_synth_fill')
            device.launch_count += 1
        return out[:n_bytes]

    def records_from(self, lo, hi):
        """(k_lo, k_hi): the records whose leading newline -- stream byte pos0 - 1, the virtual sentinel for record 0 --
        lies in [lo, hi): the records a shard holding these bytes owns."""
        q = torch.tensor([int(lo), int(hi)], dtype=torch.int64, device=self.dev)
        # pos0 - 1 >= lo  <=>  pos0 > lo ; record 0 (pos0 = 0, sentinel) belongs to the shard that starts at 0
        idx = torch.searchsorted(self.off[:self.n], q, right=True).tolist()
        k_lo = 0 if lo <= 0 else idx[0]
        k_hi = idx[1] if hi > 0 else 0
        return k_lo, k_hi

    def truth(self, k_lo, k_hi):
        """True offset rows [pos0..pos5] (absolute stream offsets) of records k_lo .. k_hi - 1."""
        m = k_hi - k_lo
        with torch.cuda.device(self.dev):
            meta = torch.empty((max(m, 1), 4), dtype=torch.int32, device=self.dev)
            if m:
                _lib.check(_lib.lib().fqb_synth_meta(self.kind, self.seed, k_lo, m, self._qt(), meta.data_ptr(), None,
                                                     _stream()), 'fqb_synth_meta')
                device.launch_count += 1
            meta = meta[:m].to(torch.int64)
            p0 = self.off[k_lo:k_hi]
            p1 = p0 + meta[:, 0]
            p2 = p1 + 1
            p3 = p2 + meta[:, 2]
            p4 = p3 + 1 + meta[:, 3] + 1
            p5 = p4 + meta[:, 2]
            return torch.stack([p0, p1, p2, p3, p4, p5], dim=1)

    def mismatches(self, rows, k_first, chunk=1 << 22):
        """Number of rows of `rows` (int64 [m,6], the parser's table for records k_first ..) that differ from the
        truth, compared in chunks (the truth of 64 GiB is 9 GB)."""
        bad = 0
        for a in range(0, rows.shape[0], chunk):
            b = min(rows.shape[0], a + chunk)
            want = self.truth(k_first + a, k_first + b)
            bad += int((rows[a:b] != want).any(dim=1).sum().item())
        return bad


def host_record(kind, k, stream_offset, seed=None):
    """Bytes and (hl, rl, sb, pl) of ONE record, computed on the host by the library (no device involved)."""
    L = _lib.lib()
    qt = ont_qtable() if kind == 'ont' else None
    qp = qt.ctypes.data_as(ctypes.c_void_p) if qt is not None else None
    seed = SEEDS[kind] if seed is None else int(seed)
    meta = (ctypes.c_int32 * 4)()
    n = L.fqb_synth_host_record(KINDS[kind], seed, int(k), int(stream_offset), qp, None, 0, meta)
    buf = (ctypes.c_uint8 * n)()
    L.fqb_synth_host_record(KINDS[kind], seed, int(k), int(stream_offset), qp, buf, n, meta)
    return bytes(buf), tuple(meta)
