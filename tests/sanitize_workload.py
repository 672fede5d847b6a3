"""A small pass over every kernel family, meant to run under compute-sanitizer (tools/gpu_sanitize.sh):
fast path (aligned / unaligned buffers, fused and byte-exact Phred decode, dense lists), general path, FASTA,
the consumers of the table, the sharded protocol with local exchanges (plain, publish/wait, decode, general).
Results are checked against the oracle as in the tests; sizes are kept to a few MiB because the tools slow the
kernels down by one to two orders of magnitude."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))  # fqgen, algo_model

import numpy as np  # noqa: E402


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    import fastqandfurious_b200 as fq
    from fastqandfurious_b200 import shard
    import fqgen
    import oracle

    def dev(data, offset=0):
        a = np.frombuffer(bytes(data), dtype=np.uint8)
        t = torch.full((len(a) + offset + 32,), 10, dtype=torch.uint8, device='cuda')
        t[offset:offset + len(a)].copy_(torch.from_numpy(a.copy()))
        return t[offset:offset + len(a)]

    def check(data, offset=0, **kw):
        want, st, tail, resume = oracle.parse_chain(b'\n' + bytes(data), 0, -1)
        res = fq.parse_buffer(dev(data, offset), **kw)
        assert np.array_equal(res.table.cpu().numpy(), want) and res.tail_status == st
        if res.qual is not None:
            got = res.table.cpu().numpy()
            q = res.qual.cpu().numpy()
            assert np.array_equal(np.concatenate([q[r[4]:r[5]] for r in got]), oracle.decode_quals(bytes(data), got))
        return res

    rng = random.Random(5)
    fixed = fqgen.fixed_records_np(9000).tobytes()               # 3 MB, ~185 tiles
    multi = fqgen.variable_records_np(3000, 3, 'multiline').tobytes()
    ont = fqgen.variable_records_np(120, 3, 'ont').tobytes()
    n = 0
    for data in (fixed, ont):
        for off in (0, 7):
            for cfg in (0, 1, 3):
                assert check(data, off, cfg=cfg, decode_quality=True).path == 1
                n += 1
    assert check(multi, 3, decode_quality=True).path == 2
    # the speculative pass in both forms (a warp per chunk: the default; a CTA per chunk), and on the 32 KiB scan geometry
    r = check(multi, 5, force_general=True)
    assert r.path == 2 and r.spec
    assert check(multi, 0, force_general=True, spec='v1').spec
    assert check(multi, 9, force_general=True, cfg=2).spec
    assert check(fixed[:200000], 0, force_general=True).path == 2
    check(b'@a\nA\n+\nI\n' * 20000, 1)                           # dense lists (FQB_FLAG_DENSE retry)
    for data in fqgen.corpus(9000, 25):
        check(data, rng.randrange(16))
    n += 30
    # FASTA
    fasta = b''.join(b'>r%d d\n' % k + fqgen._wrap(bytes(rng.choice(b'ACGT') for _ in range(rng.randint(0, 400))), 60) + b'\n'
                     for k in range(3000)) + b'>\n' * 50000 + b'>y\nAC\n>z\n'
    want, st, tail, resume = oracle.fasta_chain(b'\n' + fasta, 0, -1)
    for cfg in (0, 1):  # both scan geometries of the FASTA parse
        res = fq.parse_fasta_buffer(dev(fasta, 5), cfg=cfg)
        assert np.array_equal(res.table.cpu().numpy(), want) and res.tail_status == st
    # consumers
    d = dev(fixed)
    table = fq.parse_buffer(d).table
    tn = table.cpu().numpy()
    for field, fid in (('sequence', 1), ('quality', 2), ('header', 0)):
        got, offs = fq.gather_fields(d, table, field)
        w, wo = oracle.gather_fields(fixed, tn, fid)
        assert np.array_equal(got.cpu().numpy(), w) and np.array_equal(offs.cpu().numpy(), wo)
    for dd, dt_ in ((d, table), (dev(multi, 3), None), (dev(ont, 1), None)):
        tt = fq.parse_buffer(dd).table if dt_ is None else dt_
        src = fixed if dt_ is not None else (multi if dd.numel() == len(multi) else ont)
        packed, offs, nb, no = fq.pack_2bit(dd, tt)
        wp, wo, wnb, wno = oracle.pack_2bit(src, tt.cpu().numpy())
        assert np.array_equal(packed.cpu().numpy(), wp) and np.array_equal(nb.cpu().numpy(), wnb)
    sel = fq.select_by_length(table, 100, 200)
    assert len(sel) == len(tn)
    sums = fq.field_sums(d, table, 'quality', add=-33)
    q = np.frombuffer(fixed, np.uint8).reshape(-1, 337)[:len(tn), 186:336].astype(np.int64) - 33
    assert np.array_equal(sums.cpu().numpy(), q.sum(1))
    # sharded protocol, exchanges replaced by local copies
    want = oracle.parse_chain(b'\n' + fixed, 0, -1)[0]
    for fused, tail in ((False, False), (True, False), (True, True), (False, True)):
        quals = []
        rows, last = shard.parse_shards_local(d, [1000001, 2000002], halo_bytes=4096, fused=fused, epoch=3,
                                              quals_out=quals, tail=tail)
        assert np.array_equal(torch.cat(rows).cpu().numpy(), want)
        for r, (off, qq) in zip(rows, quals):
            r = r.cpu().numpy()
            qq = qq.cpu().numpy()
            assert np.array_equal(np.concatenate([qq[a - off:b - off] for a, b in zip(r[:, 4], r[:, 5])]),
                                  oracle.decode_quals(fixed, r))
    dm = dev(multi)
    rows, results = shard.parse_shards_local_general(dm, [len(multi) // 3, 2 * len(multi) // 3], halo_bytes=8192)
    wantm = oracle.parse_chain(b'\n' + multi, 0, -1)[0]
    assert all(r.error == 0 for r in results)
    assert np.array_equal(torch.cat(rows).cpu().numpy(), wantm)
    t = torch.arange(-100, 100, dtype=torch.int8, device='cuda')
    fq.device.arrayadd_b_(t, -33)
    torch.cuda.synchronize()
    print('sanitize workload ok (%d fast-path parses)' % n)


if __name__ == '__main__':
    main()
