"""ctypes binding of libfqb200.so (C ABI declared in include/fqb200.h).

The shared library holds the sm_100a kernels; it is built in-tree by ``__graft_entry__.build()``
(``nvcc -gencode arch=compute_100a,code=sm_100a``).  There is NO fallback: if the library is
missing every entry point of this package raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FQB200_LIB: another build of the same sources (tools/gpu_sanitize.sh loads the racecheck variant)
LIBPATH = os.environ.get('FQB200_LIB') or os.path.join(_HERE, 'libfqb200.so')

# status codes of the reference (src/_fastqandfurious.c:7-15, src/fastqandfurious.py:19-27)
INVALID = -1
MISSING_SEQHEADER_BEGIN = POS_HEAD_BEG = 0
MISSING_SEQHEADER_END = POS_HEAD_END = 1
MISSING_SEQ_BEG = POS_SEQ_BEG = 2
MISSING_SEQ_END = POS_SEQ_END = 3
MISSING_QUAL_BEGIN = POS_QUAL_BEG = 4
MISSING_QUAL_END = POS_QUAL_END = 5
COMPLETE = 6
MISSING_QUALHEADER_END = 7

ERR_OK, ERR_CAPACITY, ERR_WORKSPACE, ERR_TOO_MANY_LINES, ERR_DENSE, ERR_HALO, ERR_SHARD_GENERAL, ERR_PEER, ERR_OVERRUN = 0, 1, 2, 3, 4, 5, 6, 7, 8
SYNTH_ILLUMINA, SYNTH_ONT, SYNTH_MULTILINE = 0, 1, 2
PATH_FAST4, PATH_GENERAL = 1, 2
FLAG_FORCE_GENERAL, FLAG_FAST_ONLY, FLAG_DENSE, FLAG_NO_SPEC, FLAG_SPEC_ONLY, FLAG_SPEC_V1 = 1, 2, 4, 8, 16, 32
FLAG_SHARD_TAIL = 0x10000


def FLAG_CFG(i):
    return (int(i) & 15) << 8


class FqbResult(ctypes.Structure):
    """struct fqb_result of include/fqb200.h (128 bytes)."""
    _fields_ = [('n_records', ctypes.c_int64), ('resume_offset', ctypes.c_int64),
                ('tail_pos', ctypes.c_int64 * 6), ('tail_status', ctypes.c_int32), ('path', ctypes.c_int32),
                ('error', ctypes.c_int32), ('need_general', ctypes.c_int32), ('n_lines', ctypes.c_int64),
                ('first_bad', ctypes.c_int64), ('reserved', ctypes.c_int64 * 4)]


assert ctypes.sizeof(FqbResult) == 128

SYMBOLS = ('fqb_workspace_bytes', 'fqb_parse', 'fqb_shard_scan', 'fqb_shard_emit', 'fqb_shard_scan_publish', 'fqb_shard_scan_decode', 'fqb_shard_scan_publish_ready', 'fqb_shard_emit_wait', 'fqb_shard_pull_halo', 'fqb_shard_signal_ready', 'fqb_shard_wait_ready', 'fqb_shard_general', 'fqb_sum_u64_ptrs', 'fqb_arrayadd_b', 'fqb_arrayadd_q', 'fqb_synth_fixed',
           'fqb_kernel_info', 'fqb_version', 'fqb_profile_enable', 'fqb_profile_read', 'fqb_field_lengths', 'fqb_length_flags',
           'fqb_scan_workspace_bytes', 'fqb_exclusive_scan', 'fqb_compact_indices', 'fqb_gather_fields', 'fqb_field_sums', 'fqb_pack_2bit', 'fqb_fasta_workspace_bytes', 'fqb_parse_fasta',
           'fqb_synth_meta', 'fqb_synth_fill', 'fqb_synth_host_record')

_lib = None


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """The loaded library; raises ExtensionMissing when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise ExtensionMissing('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU fallback)' % LIBPATH)
    L = ctypes.CDLL(LIBPATH)
    i32, i64, u32, u64, p, sz = (ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64,
                                 ctypes.c_void_p, ctypes.c_size_t)
    L.fqb_workspace_bytes.argtypes = [i64, i64, u32]
    L.fqb_workspace_bytes.restype = sz
    L.fqb_parse.argtypes = [p, i64, i32, i64, p, i64, p, i32, p, p, sz, i64, u32, p]
    L.fqb_parse.restype = ctypes.c_int
    L.fqb_arrayadd_b.argtypes = [p, i64, i32, p]
    L.fqb_arrayadd_b.restype = ctypes.c_int
    L.fqb_arrayadd_q.argtypes = [p, i64, i64, p]
    L.fqb_arrayadd_q.restype = ctypes.c_int
    L.fqb_synth_fixed.argtypes = [p, i64, i64, i32, i32, u64, p]
    L.fqb_shard_scan.argtypes = [p, i64, i64, i32, p, p, sz, u32, p]
    L.fqb_shard_scan.restype = ctypes.c_int
    L.fqb_shard_emit.argtypes = [p, i64, i64, i32, i32, i64, p, p, i64, p, p, sz, u32, p]
    L.fqb_shard_emit.restype = ctypes.c_int
    L.fqb_shard_scan_publish.argtypes = [p, i64, i64, i32, p, p, i32, u64, p, sz, u32, p]
    L.fqb_shard_scan_publish.restype = ctypes.c_int
    L.fqb_shard_scan_decode.argtypes = [p, i64, i64, i32, p, p, i32, u64, p, i32, p, sz, u32, p]
    L.fqb_shard_scan_decode.restype = ctypes.c_int
    L.fqb_shard_scan_publish_ready.argtypes = [p, i64, i64, i32, p, p, i32, u64, p, u64, p, i32, p, sz, u32, p]
    L.fqb_shard_scan_publish_ready.restype = ctypes.c_int
    L.fqb_shard_emit_wait.argtypes = [p, i64, i64, i32, i32, i64, p, i32, u64, p, i64, p, p, sz, u32, p]
    L.fqb_shard_emit_wait.restype = ctypes.c_int
    L.fqb_shard_pull_halo.argtypes = [p, p, i64, p, p, u64, p, p]
    L.fqb_shard_pull_halo.restype = ctypes.c_int
    L.fqb_shard_general.argtypes = [p, i64, i64, i32, i32, i32, i64, p, p, p, u64, u64, p, i64, p, p, sz, i64, u32, p]
    L.fqb_shard_general.restype = ctypes.c_int
    L.fqb_shard_signal_ready.argtypes = [p, u64, p]
    L.fqb_shard_signal_ready.restype = ctypes.c_int
    L.fqb_shard_wait_ready.argtypes = [p, u64, p, p]
    L.fqb_shard_wait_ready.restype = ctypes.c_int
    L.fqb_sum_u64_ptrs.argtypes = [p, i32, p, p]
    L.fqb_sum_u64_ptrs.restype = ctypes.c_int
    L.fqb_synth_fixed.restype = ctypes.c_int
    L.fqb_kernel_info.argtypes = [i32, p, p, p, p]
    L.fqb_kernel_info.restype = ctypes.c_int
    L.fqb_profile_enable.argtypes = [i32]
    L.fqb_profile_enable.restype = ctypes.c_int
    L.fqb_profile_read.argtypes = [p, p]
    L.fqb_profile_read.restype = ctypes.c_int
    L.fqb_field_lengths.argtypes = [p, i64, p, i64, i32, p, p, p]
    L.fqb_field_lengths.restype = ctypes.c_int
    L.fqb_length_flags.argtypes = [p, i64, i32, i64, i64, p, p, p]
    L.fqb_length_flags.restype = ctypes.c_int
    L.fqb_scan_workspace_bytes.argtypes = [i64]
    L.fqb_scan_workspace_bytes.restype = sz
    L.fqb_exclusive_scan.argtypes = [p, i64, p, p, sz, p]
    L.fqb_exclusive_scan.restype = ctypes.c_int
    L.fqb_compact_indices.argtypes = [p, i64, p, p]
    L.fqb_compact_indices.restype = ctypes.c_int
    L.fqb_gather_fields.argtypes = [p, i64, i64, p, i64, p, i64, i32, p, p, i32, p, p]
    L.fqb_gather_fields.restype = ctypes.c_int
    L.fqb_field_sums.argtypes = [p, i64, i64, p, i64, p, i64, i32, i32, p, p, p]
    L.fqb_field_sums.restype = ctypes.c_int
    L.fqb_pack_2bit.argtypes = [p, i64, i64, p, i64, p, i64, p, p, p, p, p, p]
    L.fqb_pack_2bit.restype = ctypes.c_int
    L.fqb_fasta_workspace_bytes.argtypes = [i64, i64, u32]
    L.fqb_fasta_workspace_bytes.restype = sz
    L.fqb_parse_fasta.argtypes = [p, i64, i32, i64, p, i64, p, p, sz, i64, u32, p]
    L.fqb_parse_fasta.restype = ctypes.c_int
    L.fqb_synth_meta.argtypes = [i32, u64, i64, i64, p, p, p, p]
    L.fqb_synth_meta.restype = ctypes.c_int
    L.fqb_synth_fill.argtypes = [i32, u64, i64, i64, p, p, p, i64, i64, p]
    L.fqb_synth_fill.restype = ctypes.c_int
    L.fqb_synth_host_record.argtypes = [i32, u64, i64, i64, p, p, i64, p]
    L.fqb_synth_host_record.restype = i64
    L.fqb_version.argtypes = []
    L.fqb_version.restype = ctypes.c_char_p
    _lib = L
    return L


def check(code, what):
    if code != 0:
        raise RuntimeError('%s failed: cudaError %d' % (what, code))
