"""Step times (CUDA events) and scan-kernel times (fqb_profile_*) of the device paths for whatever build FQB200_LIB
points at -- A/B runs of library variants built from the same tree:  FQB200_LIB=variants/x.so python tools/ab_paths.py
[fast dec ont illumina multiline fasta] (1 GiB inputs, generated on the device)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'fastq-and-furious_b200'), os.path.join(ROOT, 'tests')]
import torch  # noqa: E402

import fastqandfurious_b200 as fq  # noqa: E402
from fastqandfurious_b200 import _lib, device, shard  # noqa: E402

L = _lib.lib()
want = sys.argv[1:] or ['fast', 'dec', 'ont', 'illumina', 'multiline']
GIB = 1 << 30
tag = os.path.basename(_lib.LIBPATH)


def timed(name, fn, nbytes, steps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms0 = e0.elapsed_time(e1) / steps  # no events between the kernels of a step
    _lib.check(L.fqb_profile_enable(1), 'profile')
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    _lib.check(L.fqb_profile_read(ctypes.byref(tot), ctypes.byref(cnt)), 'read')
    L.fqb_profile_enable(0)
    ms = e0.elapsed_time(e1) / steps
    sms = tot.value / max(1, cnt.value)
    print('%-24s %-16s step %.4f ms = %7.1f GB/s (with scan events %.4f) | scan kernel %.4f ms = %7.1f GB/s' % (
        tag, name, ms0, nbytes / ms0 / 1e6, ms, sms, nbytes / sms / 1e6 if sms else 0), flush=True)


result = torch.empty(16, dtype=torch.int64, device='cuda')
if 'fast' in want or 'dec' in want:
    n = GIB // 337
    buf = fq.synth_fixed(n)
    table = torch.empty((n + 64, 6), dtype=torch.int64, device='cuda')
    if 'fast' in want:
        timed('fixed150', lambda: device.parse_raw(buf, 1, -1, table, None, 0, result, _lib.FLAG_FAST_ONLY), buf.numel(), 100)
        res = device.read_result(result)
        assert res.error == 0 and res.n_records == n - 1 and not res.need_general, (res.error, res.n_records)
    if 'dec' in want:
        qual = torch.full((buf.numel(),), 0x5a, dtype=torch.int8, device='cuda')
        timed('fixed150+decode', lambda: device.parse_raw(buf, 1, -1, table, qual, -33, result, _lib.FLAG_FAST_ONLY),
              buf.numel())
        print('   mirror bytes left untouched: %.3f, ranges rewritten: %d' % (
            float((qual == 0x5a).float().mean()), device.read_result(result).reserved[2]), flush=True)
        if hasattr(_lib, 'FLAG_DEC_WHOLE'):
            timed('fixed150+dec(whole)', lambda: device.parse_raw(buf, 1, -1, table, qual, -33, result,
                                                                  _lib.FLAG_FAST_ONLY | _lib.FLAG_DEC_WHOLE), buf.numel())
        rows = table[:65536]
        idx = rows[:, 4].unsqueeze(1) + torch.arange(150, device='cuda')
        assert torch.equal(qual[idx.reshape(-1)], (buf[idx.reshape(-1)].to(torch.int16) - 33).to(torch.int8))
        del qual
    del buf, table
    device._ws_cache.clear()
    torch.cuda.empty_cache()
for name, kind in (('ont', 'ont'), ('illumina', 'illumina'), ('multiline', 'multiline')):
    if name not in want:
        continue
    for cfg in [int(c) for c in os.environ.get('AB_CFGS', '0').split()]:  # scan kernel geometries (FQB_FLAG_CFG)
        job = shard.SynthJob(kind, GIB, 0, 1, 'cuda', cfg=cfg)
        job.prepare()
        job.step()
        torch.cuda.synchronize()
        ok = job.verify_local()
        timed(name + (' (exact)' if job.exact else '') + (' cfg %d' % cfg if cfg else ''), job.step, job.global_bytes())
        print('   rows verified', ok, flush=True)
        job.free()
if 'fasta' in want:  # records of 300 bases wrapped at 60 columns, 1 GiB (the bench's fasta_1g)
    import numpy as np
    rng = np.random.default_rng(6)
    nrec = 190000
    rec = bytearray()
    seqs = rng.choice(np.frombuffer(b'ACGT', dtype=np.uint8), size=(nrec, 5, 60))
    for k in range(nrec):
        rec += b'>read%07d sample\n' % k
        rec += b'\n'.join(bytes(row) for row in seqs[k]) + b'\n'
    base = np.frombuffer(bytes(rec), dtype=np.uint8)
    d = torch.from_numpy(base.copy()).cuda().repeat(max(1, GIB // len(base)))
    fcfg = int(os.environ.get('FASTA_CFG', '0'))  # FASTA scan geometry: 1 = 16 KiB per iteration, else 32 KiB (the default)
    res = device.parse_fasta_buffer(d, cfg=fcfg)
    want_n = nrec * max(1, GIB // len(base)) - 1
    assert res.n == want_n, (res.n, want_n)
    t = res.table[:4096].cpu().numpy()
    step = len(rec) // nrec
    assert (np.diff(t[:, 0]) == step).all() and (np.diff(t[:, 3]) == step).all() and (t[:, 2] - t[:, 1] == 1).all()
    ml, cap4 = int(res.n_lines) + 64, int(res.n) + 64
    tab4 = torch.empty((cap4, 4), dtype=torch.int64, device='cuda')
    fl = _lib.FLAG_CFG(fcfg)
    ws = torch.empty(L.fqb_fasta_workspace_bytes(d.numel(), ml, fl) + 256, dtype=torch.uint8, device='cuda')
    timed('fasta' + (' cfg %d' % fcfg if fcfg else ''),
          lambda: _lib.check(L.fqb_parse_fasta(d.data_ptr(), d.numel(), 1, -1, tab4.data_ptr(), cap4, result.data_ptr(),
                                               ws.data_ptr(), ws.numel(), ml, fl, device._stream()), 'fqb_parse_fasta'),
          d.numel())
    assert torch.equal(tab4[:res.n], res.table)
