"""Device-side consumers of the offset table: what users of the reference do per record
inside an ``entryfunc``, done for the whole table on the GPU.

* ``field_lengths`` / ``select_by_length`` -- the length filter of doc/user-guide.rst:153-180
  (``posarray[3] - posarray[2] < LENGTH_THRESHOLD``), as row indices instead of ``None`` results;
* ``gather_fields`` -- index replay (src/demo/benchmark.py:47-83): header / sequence / quality bytes of the
  selected rows, packed contiguously, optionally Phred-decoded on the way (``arrayadd_b(q, -33)``,
  src/demo/benchmark.py:161-163);
* ``field_sums`` -- per-record sum of the decoded quality bytes (mean quality = sum / length);
* ``pack_2bit`` -- the sequences at 2 bits per base (records with other letters are counted, not lost);
* ``write_index`` / ``read_index`` -- the on-disk index of src/demo/benchmark.py:268-287
  (``array('q').tofile`` of the 6 positions per record).

Fields: ``'header'`` = buf[pos0+1:pos1], ``'sequence'`` = buf[pos2:pos3], ``'quality'`` = buf[pos4:pos5]
(entryfunc, src/fastqandfurious.py:161-171).  All tensors live on the GPU; there is no CPU fallback."""
import ctypes

import numpy as np
import torch

from . import _lib, device

FIELDS = {'header': 0, 'sequence': 1, 'quality': 2, 0: 0, 1: 1, 2: 2}


def _field(field):
    try:
        return FIELDS[field]
    except KeyError:
        raise ValueError("field must be 'header', 'sequence' or 'quality'")


def _table(table):
    device._require_cuda(table, 'table')
    if table.dtype != torch.int64 or table.dim() != 2 or table.shape[1] != 6:
        raise TypeError('table must be an int64 [n,6] CUDA tensor')
    return table


def _sel(sel, table):
    if sel is None:
        return None, table.shape[0], None
    device._require_cuda(sel, 'sel')
    if sel.dtype != torch.int64 or sel.dim() != 1 or sel.device != table.device:
        raise TypeError('sel must be a 1-D int64 CUDA tensor of row indices on the table\'s device')
    return sel, sel.numel(), sel.data_ptr() if sel.numel() else None


def _check_status(status, what):
    code = int(status.item())
    if code & 1:
        raise IndexError('%s: a selected row index is outside the table' % what)
    if code & 2:
        raise ValueError('%s: a row has a reversed span or points outside the buffer' % what)


def field_lengths(table, field='sequence', sel=None):
    """int64 [n_sel]: length of `field` for every (selected) row."""
    table = _table(table)
    sel, n_sel, sel_ptr = _sel(sel, table)
    with torch.cuda.device(table.device):
        out = torch.empty(n_sel, dtype=torch.int64, device=table.device)
        status = torch.zeros(1, dtype=torch.int32, device=table.device)
        _lib.check(_lib.lib().fqb_field_lengths(table.data_ptr() if table.numel() else None, table.shape[0], sel_ptr,
                                                n_sel, _field(field), out.data_ptr() if n_sel else None,
                                                status.data_ptr(), device._stream()), 'fqb_field_lengths')
        device.launch_count += 1
        _check_status(status, 'field_lengths')
    return out


def exclusive_scan(values):
    """int64 [n+1]: exclusive prefix sums of an int64 CUDA vector (out[n] = total)."""
    device._require_cuda(values, 'values')
    if values.dtype != torch.int64 or values.dim() != 1:
        raise TypeError('values must be a 1-D int64 CUDA tensor')
    n = values.numel()
    L = _lib.lib()
    with torch.cuda.device(values.device):
        out = torch.empty(n + 1, dtype=torch.int64, device=values.device)
        need = L.fqb_scan_workspace_bytes(n)
        ws = torch.empty(max(need, 8), dtype=torch.uint8, device=values.device)
        _lib.check(L.fqb_exclusive_scan(values.data_ptr() if n else None, n, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                        device._stream()), 'fqb_exclusive_scan')
        device.launch_count += 3 if n else 1
    return out


def select_by_length(table, min_len=0, max_len=2 ** 62, field='sequence'):
    """int64 [m]: indices (ascending) of the rows whose `field` length lies in [min_len, max_len]."""
    table = _table(table)
    n = table.shape[0]
    L = _lib.lib()
    with torch.cuda.device(table.device):
        flags = torch.empty(n, dtype=torch.int64, device=table.device)
        status = torch.zeros(1, dtype=torch.int32, device=table.device)
        _lib.check(L.fqb_length_flags(table.data_ptr() if n else None, n, _field(field), int(min_len), int(max_len),
                                      flags.data_ptr() if n else None, status.data_ptr(), device._stream()),
                   'fqb_length_flags')
        excl = exclusive_scan(flags)
        m = int(excl[-1].item())
        idx = torch.empty(m, dtype=torch.int64, device=table.device)
        if m:
            _lib.check(L.fqb_compact_indices(excl.data_ptr(), n, idx.data_ptr(), device._stream()), 'fqb_compact_indices')
        device.launch_count += 2
        _check_status(status, 'select_by_length')
    return idx


def gather_fields(buf, table, field='sequence', sel=None, add=0, table_base=0):
    """(out uint8 [total], offsets int64 [n_sel+1]): the `field` bytes of the (selected) rows packed back to
    back, record i at out[offsets[i]:offsets[i+1]]; `add` is added to every byte with int8 wrap (-33 on
    'quality' = Phred decode).  Table positions index the stream; buf[0] is stream position `table_base`."""
    device._require_cuda(buf, 'buf')
    if buf.dtype != torch.uint8 or buf.dim() != 1:
        raise TypeError('buf must be a 1-D uint8 CUDA tensor')
    table = _table(table)
    if table.device != buf.device:
        raise ValueError('buf and table must live on the same device')
    sel, n_sel, sel_ptr = _sel(sel, table)
    L = _lib.lib()
    with torch.cuda.device(buf.device):
        lens = field_lengths(table, field, sel)
        offsets = exclusive_scan(lens)
        total = int(offsets[-1].item())
        out = torch.empty(total, dtype=torch.uint8, device=buf.device)
        status = torch.zeros(1, dtype=torch.int32, device=buf.device)
        _lib.check(L.fqb_gather_fields(buf.data_ptr() if buf.numel() else None, buf.numel(), int(table_base),
                                       table.data_ptr() if table.numel() else None, table.shape[0], sel_ptr, n_sel,
                                       _field(field), offsets.data_ptr(), out.data_ptr() if total else None,
                                       device._short_to_byte(add), status.data_ptr(), device._stream()),
                   'fqb_gather_fields')
        device.launch_count += 1
        _check_status(status, 'gather_fields')
    return out, offsets


def field_sums(buf, table, field='quality', sel=None, add=-33, table_base=0):
    """int64 [n_sel]: per record, the sum of (int8)(byte + add) over the field's bytes."""
    device._require_cuda(buf, 'buf')
    if buf.dtype != torch.uint8 or buf.dim() != 1:
        raise TypeError('buf must be a 1-D uint8 CUDA tensor')
    table = _table(table)
    if table.device != buf.device:
        raise ValueError('buf and table must live on the same device')
    sel, n_sel, sel_ptr = _sel(sel, table)
    with torch.cuda.device(buf.device):
        out = torch.empty(n_sel, dtype=torch.int64, device=buf.device)
        status = torch.zeros(1, dtype=torch.int32, device=buf.device)
        _lib.check(_lib.lib().fqb_field_sums(buf.data_ptr() if buf.numel() else None, buf.numel(), int(table_base),
                                             table.data_ptr() if table.numel() else None, table.shape[0], sel_ptr,
                                             n_sel, _field(field), device._short_to_byte(add),
                                             out.data_ptr() if n_sel else None, status.data_ptr(), device._stream()),
                   'fqb_field_sums')
        device.launch_count += 1
        _check_status(status, 'field_sums')
    return out


def pack_2bit(buf, table, sel=None, table_base=0):
    """(packed uint8 [total], offsets int64 [n_sel+1], n_bases int64 [n_sel], n_other int64 [n_sel]): the sequences
    of the (selected) rows at 2 bits per base -- A/a = 0, C/c = 1, G/g = 2, T/t/U/u = 3, base k of record i in bits
    2(k % 4).. of packed[offsets[i] + k // 4]; newlines inside wrapped sequences are skipped.  Other bytes ('N', IUPAC
    codes) are encoded by the same bit formula and counted per record in n_other, so those records can be fetched as
    bytes with gather_fields.  Slots are whole 32-bit words (4 * ceil(field bytes / 16)), zero behind the last base."""
    device._require_cuda(buf, 'buf')
    if buf.dtype != torch.uint8 or buf.dim() != 1:
        raise TypeError('buf must be a 1-D uint8 CUDA tensor')
    table = _table(table)
    if table.device != buf.device:
        raise ValueError('buf and table must live on the same device')
    sel, n_sel, sel_ptr = _sel(sel, table)
    L = _lib.lib()
    with torch.cuda.device(buf.device):
        lens = field_lengths(table, 'sequence', sel)
        offsets = exclusive_scan(((lens + 15) // 16) * 4)
        total = int(offsets[-1].item())
        out = torch.empty(max(total, 4), dtype=torch.uint8, device=buf.device)[:total]
        n_bases = torch.empty(n_sel, dtype=torch.int64, device=buf.device)
        n_other = torch.empty(n_sel, dtype=torch.int64, device=buf.device)
        status = torch.zeros(1, dtype=torch.int32, device=buf.device)
        _lib.check(L.fqb_pack_2bit(buf.data_ptr() if buf.numel() else None, buf.numel(), int(table_base),
                                   table.data_ptr() if table.numel() else None, table.shape[0], sel_ptr, n_sel,
                                   offsets.data_ptr(), out.data_ptr() if total else None,
                                   n_bases.data_ptr() if n_sel else None, n_other.data_ptr() if n_sel else None,
                                   status.data_ptr(), device._stream()), 'fqb_pack_2bit')
        device.launch_count += 1
        _check_status(status, 'pack_2bit')
    return out, offsets, n_bases, n_other


# ---- the on-disk index (host I/O only) ------------------------------------------------------------
def write_index(table, fh):
    """Append the rows to a binary file exactly like ``pos.tofile(fh_index)`` per record
    (src/demo/benchmark.py:276-280): native-endian int64, 6 per record.  Returns the rows written."""
    rows = table.detach().cpu().numpy() if isinstance(table, torch.Tensor) else np.asarray(table)
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 6)
    fh.write(rows.tobytes())
    return len(rows)


def read_index(fh, device_=None):
    """All rows of an index file written by ``write_index`` / the reference's ``tofile`` loop, as an int64
    [n,6] array (CUDA tensor when `device_` is given).  A trailing partial record is an error, like the
    EOFError of ``posarray.fromfile(fh_index, 6)`` mid-record (src/demo/benchmark.py:60-63)."""
    raw = fh.read()
    if len(raw) % 48:
        raise EOFError('index file ends inside a record (%d stray bytes)' % (len(raw) % 48))
    rows = np.frombuffer(raw, dtype=np.int64).reshape(-1, 6)
    if device_ is None:
        return rows.copy()
    return torch.from_numpy(rows.copy()).to(device_)
