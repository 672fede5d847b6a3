"""CPU ORACLE -- test infrastructure, not product code.

ctypes front-end of ``oracle/fqoracle.c`` (a plain-C restatement of the reference's FASTQ
hot path) plus a loader for ``oracle/_ref`` (the UNMODIFIED reference compiled from
``/root/reference`` by ``oracle/Makefile``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
package ``fastqandfurious_b200`` never does.

Reference behaviour restated here: ``src/_fastqandfurious.c:25-217`` (entrypos, arrayadd_b,
arrayadd_q) and ``src/fastqandfurious.py:198-279`` (readfastq_iter end-of-stream rules).
"""
import ctypes
import importlib.machinery
import importlib.util
import os
import subprocess
import sys
from array import array

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, 'libfqoracle.so')

INVALID = -1
MISSING_SEQHEADER_BEGIN = 0
MISSING_SEQHEADER_END = 1
MISSING_SEQ_BEG = 2
MISSING_SEQ_END = 3
MISSING_QUAL_BEGIN = 4
MISSING_QUAL_END = 5
COMPLETE = 6
MISSING_QUALHEADER_END = 7

ERR_OK = 0
ERR_INCOMPLETE_FINAL_QUAL = 1
ERR_INCOMPLETE_ENTRY = 2
ERR_INVALID_ENTRY = 3

_lib = None


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIBPATH) or \
            os.path.getmtime(_LIBPATH) < os.path.getmtime(os.path.join(_HERE, 'fqoracle.c')):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'oracle'], stdout=sys.stderr)
    if os.path.exists('/root/reference/src/_fastqandfurious.c'):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'ref'], stdout=sys.stderr)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            build()
        L = ctypes.CDLL(_LIBPATH)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        L.fqo_entrypos.argtypes = [p, i64, i64, p]
        L.fqo_entrypos.restype = ctypes.c_int
        L.fqo_entrypos_py.argtypes = [p, i64, i64, p]
        L.fqo_entrypos_py.restype = ctypes.c_int
        L.fqo_arrayadd_b.argtypes = [p, i64, ctypes.c_int]
        L.fqo_arrayadd_b.restype = None
        L.fqo_arrayadd_q.argtypes = [p, i64, i64]
        L.fqo_arrayadd_q.restype = None
        L.fqo_parse_chain.argtypes = [p, i64, i64, i64, p, i64, p, p, p]
        L.fqo_parse_chain.restype = i64
        L.fqo_readfastq.argtypes = [p, i64, i64, p, i64, p, p]
        L.fqo_readfastq.restype = i64
        L.fqo_decode_quals.argtypes = [p, p, i64, i64, ctypes.c_int, p]
        L.fqo_decode_quals.restype = i64
        L.fqo_entrypos_fasta.argtypes = [p, i64, i64, p]
        L.fqo_entrypos_fasta.restype = ctypes.c_int
        L.fqo_fasta_chain.argtypes = [p, i64, i64, i64, p, i64, p, p, p]
        L.fqo_fasta_chain.restype = i64
        _lib = L
    return _lib


def _as_u8(blob):
    a = np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob
    return np.ascontiguousarray(a, dtype=np.uint8)


def entrypos(blob, offset, posbuffer):
    """C-extension semantics (src/_fastqandfurious.c:25-153)."""
    a = _as_u8(blob)
    pos = np.empty(6, dtype=np.int64)
    st = lib().fqo_entrypos(a.ctypes.data, a.size, offset, pos.ctypes.data)
    for i in range(6):
        posbuffer[i] = int(pos[i])
    return st


def entrypos_py(blob, offset, posbuffer):
    """Pure-Python semantics (src/fastqandfurious.py:39-100); posbuffer is NOT reset."""
    a = _as_u8(blob)
    pos = np.array([posbuffer[i] for i in range(6)], dtype=np.int64)
    st = lib().fqo_entrypos_py(a.ctypes.data, a.size, offset, pos.ctypes.data)
    for i in range(6):
        posbuffer[i] = int(pos[i])
    return st


def arrayadd_b(a, value):
    """int8 in-place add (src/_fastqandfurious.c:161-185); `a` is a writable int8 ndarray."""
    assert a.dtype == np.int8 and a.flags.c_contiguous
    lib().fqo_arrayadd_b(a.ctypes.data, a.size, int(value))


def arrayadd_q(a, value):
    """int64 in-place add (src/_fastqandfurious.c:193-217)."""
    assert a.dtype == np.int64 and a.flags.c_contiguous
    lib().fqo_arrayadd_q(a.ctypes.data, a.size, int(value))


def parse_chain(blob, offset=0, goff=0, cap=None):
    """Walk the entrypos chain over one blob.

    Returns (table[n,6] int64 = pos + goff, tail_status, tail_pos[6], resume_offset)."""
    a = _as_u8(blob)
    if cap is None:
        cap = a.size // 7 + 2
    table = np.empty((cap, 6), dtype=np.int64)
    st = ctypes.c_int32(0)
    tail = np.empty(6, dtype=np.int64)
    resume = ctypes.c_int64(0)
    n = lib().fqo_parse_chain(a.ctypes.data, a.size, offset, goff, table.ctypes.data, cap,
                              ctypes.byref(st), tail.ctypes.data, ctypes.byref(resume))
    assert n <= cap
    return table[:n].copy(), st.value, tail, resume.value


def readfastq(data, cap=None):
    """readfastq_iter(BytesIO(data), ANY fbufsize, entryfunc_abspos, C entrypos) as one call.

    `data` is the raw stream (no sentinel).  Returns (table[n,6] absolute offsets, err, err_byte)."""
    blob = np.concatenate([np.array([10], dtype=np.uint8), _as_u8(data)])
    if cap is None:
        cap = blob.size // 7 + 2
    table = np.empty((cap, 6), dtype=np.int64)
    err = ctypes.c_int32(0)
    err_byte = ctypes.c_int64(0)
    n = lib().fqo_readfastq(blob.ctypes.data, blob.size, -1, table.ctypes.data, cap,
                            ctypes.byref(err), ctypes.byref(err_byte))
    return table[:n].copy(), err.value, err_byte.value


def decode_quals(data, table, value=-33):
    """Concatenated int8 quality bytes of every record, each += value (benchmark.py:161-163).

    `data` is the raw stream and `table` holds absolute offsets into it."""
    a = _as_u8(data)
    t = np.ascontiguousarray(table, dtype=np.int64)
    total = int((t[:, 5] - t[:, 4]).sum()) if len(t) else 0
    out = np.empty(total, dtype=np.int8)
    w = lib().fqo_decode_quals(a.ctypes.data, t.ctypes.data, len(t), 0, int(value), out.ctypes.data)
    assert w == total
    return out


# ---------------------------------------------------------------------------------------------
# oracle/_ref : the compiled, unmodified reference
# ---------------------------------------------------------------------------------------------
_ref = None


def reference():
    """(fastqandfurious module, _fastqandfurious C-ext module) loaded from oracle/_ref, or None."""
    global _ref
    if _ref is not None:
        return _ref or None
    pkg = os.path.join(_HERE, '_ref', 'fastqandfurious')
    pyc = os.path.join(pkg, '__init__.pyc')
    if not os.path.exists(pyc):  # some transports drop *.pyc: same bytes under another name
        pyc = os.path.join(pkg, 'module_bytecode.bin')
    ext = [f for f in (os.listdir(pkg) if os.path.isdir(pkg) else [])
           if f.startswith('_fastqandfurious') and f.endswith('.so')]
    if not (os.path.exists(pyc) and ext):
        _ref = False
        return None
    try:
        loader = importlib.machinery.SourcelessFileLoader('fqref_fastqandfurious', pyc)
        spec = importlib.util.spec_from_loader('fqref_fastqandfurious', loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        eloader = importlib.machinery.ExtensionFileLoader('_fastqandfurious', os.path.join(pkg, ext[0]))
        espec = importlib.util.spec_from_loader('_fastqandfurious', eloader)
        cext = importlib.util.module_from_spec(espec)
        eloader.exec_module(cext)
    except Exception as e:  # wrong Python version for the .pyc, ...
        print('oracle/_ref present but not loadable: %r' % (e,), file=sys.stderr)
        _ref = False
        return None
    _ref = (mod, cext)
    return _ref


def reference_abspos(data, fbufsize=2 ** 16):
    """Run the real reference: readfastq_iter + C entrypos + entryfunc_abspos over `data`.

    Returns table[n,6] (may raise the reference's ValueError)."""
    import io
    mod, cext = reference()
    out = array('q')
    for pos in mod.readfastq_iter(io.BytesIO(bytes(data)), fbufsize,
                                  entryfunc=mod.entryfunc_abspos, entrypos=cext.entrypos):
        out.extend(pos)
    return np.frombuffer(out, dtype=np.int64).reshape(-1, 6).copy() if len(out) else \
        np.empty((0, 6), dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# consumers of the offset table, restated with the reference's own slicing recipes
# ---------------------------------------------------------------------------------------------
_FIELD_COLS = {0: (0, 1, 1), 1: (2, 3, 0), 2: (4, 5, 0)}  # field -> (begin column, end column, begin adjust)


def field_spans(table, field, sel=None):
    """(begin, end) per (selected) row: header buf[pos0+1:pos1], sequence buf[pos2:pos3], quality
    buf[pos4:pos5] -- entryfunc, src/fastqandfurious.py:161-171."""
    t = np.asarray(table, dtype=np.int64).reshape(-1, 6)
    if sel is not None:
        t = t[np.asarray(sel, dtype=np.int64)]
    cb, ce, adj = _FIELD_COLS[field]
    return t[:, cb] + adj, t[:, ce]


def field_lengths(table, field, sel=None):
    """posarray[3] - posarray[2] and friends (lengthfilter_entryfunc, doc/user-guide.rst:162-167)."""
    b, e = field_spans(table, field, sel)
    return e - b


def select_by_length(table, field, min_len, max_len):
    """Row indices a length filter keeps (doc/user-guide.rst:160-167 with both bounds)."""
    n = field_lengths(table, field)
    return np.nonzero((n >= min_len) & (n <= max_len))[0].astype(np.int64)


def gather_fields(data, table, field, sel=None, add=0, table_base=0):
    """Index replay (src/demo/benchmark.py:47-83: e = (buf[pos0:pos1], buf[pos2:pos3], buf[pos4:pos5]) per index
    row) -> (packed bytes, offsets[n+1]); `add` applied like arrayadd_b (src/_fastqandfurious.c:180-182)."""
    a = _as_u8(data)
    b, e = field_spans(table, field, sel)
    parts = [a[int(x) - table_base:int(y) - table_base] for x, y in zip(b, e)]
    out = np.concatenate(parts) if parts else np.empty(0, dtype=np.uint8)
    out = (out.astype(np.int16) + (int(add) & 0xff)).astype(np.uint8)
    offsets = np.zeros(len(parts) + 1, dtype=np.int64)
    if parts:
        offsets[1:] = np.cumsum([len(p) for p in parts])
    return out, offsets


def field_sums(data, table, field, sel=None, add=-33, table_base=0):
    """Per record: sum of the int8 values array('b').frombytes(slice); arrayadd_b(a, add) would hold
    (src/demo/benchmark.py:161-163)."""
    a = _as_u8(data)
    b, e = field_spans(table, field, sel)
    out = np.zeros(len(b), dtype=np.int64)
    for i, (x, y) in enumerate(zip(b, e)):
        q = (a[int(x) - table_base:int(y) - table_base].astype(np.int16) + (int(add) & 0xff)).astype(np.uint8).view(np.int8)
        out[i] = int(q.astype(np.int64).sum())
    return out


def pack_2bit(data, table, sel=None, table_base=0):
    """2-bit packing of the sequence slices buf[pos2:pos3] (entryfunc, src/fastqandfurious.py:161-171), newlines of
    wrapped records dropped: code = ((b >> 1) & 3) ^ ((b >> 2) & 1) (A=0, C=1, G=2, T/U=3, either case), base k in bits
    2(k % 4).. of byte k // 4 of the record's slot of 4 * ceil(len(slice) / 16) bytes.  The reference has no packed
    form: this restates the layout include/fqb200.h defines.  Returns (packed, offsets, n_bases, n_other)."""
    a = _as_u8(data)
    b, e = field_spans(table, 1, sel)
    n = len(b)
    offsets = np.zeros(n + 1, dtype=np.int64)
    n_bases = np.zeros(n, dtype=np.int64)
    n_other = np.zeros(n, dtype=np.int64)
    parts = []
    acgtu = np.zeros(256, dtype=bool)
    acgtu[[ord(c) for c in 'ACGTUacgtu']] = True
    for i, (x, y) in enumerate(zip(b, e)):
        raw = a[int(x) - table_base:int(y) - table_base]
        slot = 4 * ((len(raw) + 15) // 16)
        offsets[i + 1] = offsets[i] + slot
        seq = raw[raw != 10]
        n_bases[i] = len(seq)
        n_other[i] = int((~acgtu[seq]).sum())
        codes = np.zeros(slot * 4, dtype=np.uint8)
        codes[:len(seq)] = ((seq >> 1) & 3) ^ ((seq >> 2) & 1)
        c4 = codes.reshape(-1, 4)
        parts.append((c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8))
    packed = np.concatenate(parts) if parts else np.empty(0, dtype=np.uint8)
    return packed, offsets, n_bases, n_other


# ---------------------------------------------------------------------------------------------
# FASTA (src/fastqandfurious.py:103-143)
# ---------------------------------------------------------------------------------------------
def entrypos_fasta(blob, offset, posbuffer):
    """entrypos_fasta semantics: positions assigned as they are found, posbuffer not reset."""
    a = _as_u8(blob)
    pos = np.array([posbuffer[i] for i in range(4)], dtype=np.int64)
    st = lib().fqo_entrypos_fasta(a.ctypes.data, a.size, offset, pos.ctypes.data)
    for i in range(4):
        posbuffer[i] = int(pos[i])
    return st


def fasta_chain(blob, offset=0, goff=0, cap=None):
    """Repeated entrypos_fasta calls, each starting at pos3 of the previous record.

    Returns (table[n,4] int64 = pos + goff, tail_status, tail_pos[4] (-1 = not assigned), resume_offset)."""
    a = _as_u8(blob)
    if cap is None:
        cap = a.size // 3 + 2
    table = np.empty((cap, 4), dtype=np.int64)
    st = ctypes.c_int32(0)
    tail = np.empty(4, dtype=np.int64)
    resume = ctypes.c_int64(0)
    n = lib().fqo_fasta_chain(a.ctypes.data, a.size, offset, goff, table.ctypes.data, cap, ctypes.byref(st),
                              tail.ctypes.data, ctypes.byref(resume))
    assert n <= cap
    return table[:n].copy(), st.value, tail, resume.value
