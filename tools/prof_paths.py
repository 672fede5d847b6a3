"""Every device path once inside a cudaProfilerStart/Stop window (ncu --profile-from-start off), 1 GiB inputs.

    ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/prof_paths.py [names...]

Paths: fast (150 bp fixed, fast4), dec (fused Phred mirror), ont, illumina, multiline_spec, multiline_exact,
fasta, gather, sums, pack2.  Prints the order in which the paths ran (the launch list follows that order)."""
import ctypes
import sys

sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import numpy as np
import torch

import fastqandfurious_b200 as fq
from fastqandfurious_b200 import _lib, consume, device, shard

want = sys.argv[1:] or ['fast', 'dec', 'ont', 'illumina', 'multiline_spec', 'multiline_exact', 'fasta', 'gather', 'sums',
                        'pack2']
GIB = 1 << 30
L = _lib.lib()
cudart = torch.cuda.cudart()


def window(name, fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    cudart.cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()
    print('ran', name, flush=True)


result = torch.empty(16, dtype=torch.int64, device='cuda')
if {'fast', 'dec', 'gather', 'sums', 'pack2'} & set(want):
    buf = fq.synth_fixed(GIB // 337)
    table = torch.empty((GIB // 337 + 64, 6), dtype=torch.int64, device='cuda')
    if 'fast' in want:
        window('fast', lambda: device.parse_raw(buf, 1, -1, table, None, 0, result, _lib.FLAG_FAST_ONLY))
    if 'dec' in want:
        qual = torch.empty(buf.numel(), dtype=torch.int8, device='cuda')
        window('dec', lambda: device.parse_raw(buf, 1, -1, table, qual, -33, result, _lib.FLAG_FAST_ONLY))
        del qual
    if {'gather', 'sums', 'pack2'} & set(want):
        rows = fq.parse_buffer(buf, cap=table.shape[0], table=table).table
        lens = consume.field_lengths(rows, 'sequence')
        offsets = consume.exclusive_scan(lens)
        total = int(offsets[-1].item())
        packed = torch.empty(total, dtype=torch.uint8, device='cuda')
        status = torch.zeros(1, dtype=torch.int32, device='cuda')
        sums = torch.empty(rows.shape[0], dtype=torch.int64, device='cuda')
        n = rows.shape[0]
        if 'gather' in want:
            window('gather', lambda: _lib.check(L.fqb_gather_fields(
                buf.data_ptr(), buf.numel(), 0, rows.data_ptr(), n, None, n, 1, offsets.data_ptr(), packed.data_ptr(), 0,
                status.data_ptr(), device._stream()), 'gather'))
        if 'sums' in want:
            window('sums', lambda: _lib.check(L.fqb_field_sums(
                buf.data_ptr(), buf.numel(), 0, rows.data_ptr(), n, None, n, 2, (-33) & 0xff, sums.data_ptr(),
                status.data_ptr(), device._stream()), 'sums'))
        if 'pack2' in want:
            slot_off = consume.exclusive_scan(((lens + 15) // 16) * 4)
            pk = torch.empty(int(slot_off[-1].item()), dtype=torch.uint8, device='cuda')
            nb = torch.empty(n, dtype=torch.int64, device='cuda')
            window('pack2', lambda: _lib.check(L.fqb_pack_2bit(
                buf.data_ptr(), buf.numel(), 0, rows.data_ptr(), n, None, n, slot_off.data_ptr(), pk.data_ptr(),
                nb.data_ptr(), None, status.data_ptr(), device._stream()), 'pack2'))
            del pk, nb, slot_off
        del packed, sums, lens, offsets, rows
    del buf, table
    device._ws_cache.clear()
    torch.cuda.empty_cache()

for name, kind, general, exact in (('ont', 'ont', False, False), ('illumina', 'illumina', False, False),
                                   ('multiline_spec', 'multiline', True, False),
                                   ('multiline_exact', 'multiline', True, True)):
    if name not in want:
        continue
    job = shard.SynthJob(kind, GIB, 0, 1, 'cuda', general=general)
    job.exact = exact
    job.step()
    torch.cuda.synchronize()
    res = device.read_result(job.result)
    print(name, 'records', res.n_records, 'lines', res.n_lines, 'path', res.path, 'spec', res.reserved[1], 'need_general',
          res.need_general, flush=True)
    window(name, job.step)
    job.free()

if 'fasta' in want:
    rng = np.random.default_rng(6)
    nrec = 190000
    rec = bytearray()
    seqs = rng.choice(np.frombuffer(b'ACGT', dtype=np.uint8), size=(nrec, 5, 60))
    for k in range(nrec):
        rec += b'>read%07d sample\n' % k
        rec += b'\n'.join(bytes(row) for row in seqs[k]) + b'\n'
    base = np.frombuffer(bytes(rec), dtype=np.uint8)
    d = torch.from_numpy(base.copy()).cuda().repeat(max(1, GIB // len(base)))
    res = device.parse_fasta_buffer(d)
    ml, cap4 = int(res.n_lines) + 64, int(res.n) + 64
    tab4 = torch.empty((cap4, 4), dtype=torch.int64, device='cuda')
    ws = torch.empty(L.fqb_fasta_workspace_bytes(d.numel(), ml, 0) + 256, dtype=torch.uint8, device='cuda')
    window('fasta', lambda: _lib.check(L.fqb_parse_fasta(d.data_ptr(), d.numel(), 1, -1, tab4.data_ptr(), cap4,
                                                         result.data_ptr(), ws.data_ptr(), ws.numel(), ml, 0,
                                                         device._stream()), 'fasta'))
