"""Sharding: host-side protocol on CPU (gloo, world_size 2) and the device path with the exchanges
simulated on one GPU."""
import os
import random
import socket

import numpy as np
import pytest

import fqgen


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, data, own_lens, halo, out):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fastqandfurious_b200 import shard
    plan = shard.ShardPlan(rank, world, own_lens, halo)
    plan.check()
    buf = torch.zeros(plan.own_len + plan.halo_len(), dtype=torch.uint8)
    buf[:plan.own_len] = torch.from_numpy(np.frombuffer(data, dtype=np.uint8)[plan.offset:plan.offset + plan.own_len].copy())
    got = shard.exchange_halo(buf, plan)
    own_lines = torch.tensor([int((buf[:plan.own_len] == 10).sum()) + (1 if rank == 0 else 0)], dtype=torch.int64)
    base, total = shard.line_bases(own_lines, plan)
    # what the device does with these two results, as a sequential model: the rows this shard owns
    import algo_model
    k0, rows, err = algo_model.model_shard_fast4(buf.numpy().tobytes(), plan.own_len, 1 if rank == 0 else 0, plan.is_last,
                                                 plan.offset, int(base.item()))
    out.put((rank, got, bytes(buf.numpy().tobytes()), int(base.item()), int(total.item()), k0, rows, err))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_halo_exchange_and_line_bases_gloo(oracle, world):
    """gloo run of the N>1 host logic on CPU: ring-shift halo exchange and the line-base prefix; the shards' rows
    (sequential model of the device's ownership rule fed with the exchanged bytes and bases) tile the oracle's chain."""
    import torch.multiprocessing as mp
    rng = random.Random(world)
    data = fqgen.fastq_bytes(rng, 400, read_len=(20, 60), header_len=(5, 20), long_plus=0.3, trailing_newlines=1,
                             at_plus_bias=0.3)
    cuts = [len(data) * (g + 1) // world + 7 * (g + 1) for g in range(world - 1)]
    own_lens = [b - a for a, b in zip([0] + cuts, cuts + [len(data)])]
    cut = cuts[0]
    halo = 300
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, data, own_lens, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = [0] + cuts + [len(data)]
    all_rows, next_k = [], 0
    for g, (r, got, buf, base, tot, k0, rows, err) in enumerate(res):
        last = g == world - 1
        assert r == g and got == (0 if last else halo)
        assert buf == data[bounds[g]:bounds[g + 1] + (0 if last else halo)]
        assert base == data[:bounds[g]].count(b'\n') + (1 if g else 0) and tot == data.count(b'\n') + 1
        assert err is None and k0 == next_k  # the shards' record ranges tile the stream
        next_k += len(rows)
        all_rows += rows
    want = oracle.parse_chain(b'\n' + data, 0, -1)[0]
    assert np.array_equal(np.array(all_rows, dtype=np.int64).reshape(-1, 6), want)


def test_shard_plan_bookkeeping():
    from fastqandfurious_b200 import shard
    plan = shard.ShardPlan(1, 4, [100, 50, 70, 10], halo_bytes=30)
    assert plan.offsets == [0, 100, 150, 220] and plan.total == 230 and plan.own_len == 50 and not plan.is_last
    assert [plan.halo_len(g) for g in range(4)] == [30, 30, 10, 0]
    assert [plan.send_len(g) for g in range(4)] == [0, 30, 30, 10]
    with pytest.raises(ValueError):
        shard.ShardPlan(0, 3, [100, 20, 100], halo_bytes=30).check()


@pytest.mark.gpu
def test_sharded_parse_matches_single_buffer(oracle):
    """P shards cut at arbitrary bytes, exchanges simulated by copies: rows == the single-buffer parse."""
    import torch
    import __graft_entry__ as g
    g.build()
    import fastqandfurious_b200 as fq
    from fastqandfurious_b200 import shard
    rng = random.Random(7)
    for trial in range(12):
        data = fqgen.fastq_bytes(rng, rng.randint(300, 1500), read_len=(30, 120), header_len=(5, 30), long_plus=0.3,
                                 trailing_newlines=rng.randint(0, 2), at_plus_bias=0.3)
        want, st, tail, resume = oracle.parse_chain(b'\n' + data, 0, -1)
        d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        world = rng.choice([2, 3, 4, 8])
        cuts = sorted(rng.sample(range(2000, len(data) - 2000), world - 1))
        if min(b - a for a, b in zip([0] + cuts, cuts + [len(data)])) < 1200:
            continue
        rows, last = shard.parse_shards_local(d, cuts, halo_bytes=1000, fused=bool(trial & 1), epoch=trial + 1,
                                              tail=bool(trial & 2))  # count / publish / signal in the scan's epilogue
        assert rows is not None, last.error
        got = torch.cat(rows).cpu().numpy()
        assert np.array_equal(got, want), (trial, world, cuts)
        assert last.tail_status == st
        k0s = np.cumsum([0] + [len(r) for r in rows])[:-1]
        assert last.reserved[0] == k0s[-1]
    # cuts right at / around record boundaries
    data = fqgen.fixed_records_np(3000).tobytes()
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    want = oracle.parse_chain(b'\n' + data, 0, -1)[0]
    for cut in (337 * 1000 - 1, 337 * 1000, 337 * 1000 + 1, 337 * 1000 + 33, 337 * 1000 + 184, 337 * 1000 + 186):
        for fused in (False, True):
            for tail in (False, True):
                rows, last = shard.parse_shards_local(d, [cut, cut + 337 * 900 + 5], halo_bytes=4096, fused=fused, tail=tail)
                assert np.array_equal(torch.cat(rows).cpu().numpy(), want), (cut, fused, tail)


@pytest.mark.gpu
def test_sharded_parse_with_phred_decode(oracle):
    """fqb_shard_scan_decode: every shard's mirror holds the decoded quality strings of the records it owns
    (arrayadd_b(-33) of the reference on each of them, via the oracle), also for records that end in the halo."""
    import torch
    import __graft_entry__ as g
    g.build()
    from fastqandfurious_b200 import _lib, shard
    rng = random.Random(21)
    for trial in range(8):
        data = fqgen.fastq_bytes(rng, rng.randint(800, 3000), read_len=(30, 160), header_len=(5, 30), long_plus=0.3,
                                 trailing_newlines=1, at_plus_bias=0.3)
        want = oracle.parse_chain(b'\n' + data, 0, -1)[0]
        d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        world = rng.choice([2, 3, 5])
        cuts = sorted(rng.sample(range(3000, len(data) - 3000), world - 1))
        if min(b - a for a, b in zip([0] + cuts, cuts + [len(data)])) < 1500:
            continue
        quals = []
        add = rng.choice([-33, -64, 5])
        rows, last = shard.parse_shards_local(d, cuts, halo_bytes=1200, fused=bool(trial & 1), epoch=trial + 1,
                                              quals_out=quals, qual_add=add, tail=bool(trial & 2))
        assert rows is not None, last.error
        assert np.array_equal(torch.cat(rows).cpu().numpy(), want)
        assert len(quals) == world
        for r, (off, q) in zip(rows, quals):
            r = r.cpu().numpy()
            q = q.cpu().numpy()
            got = np.concatenate([q[a - off:b - off] for a, b in zip(r[:, 4], r[:, 5])]) if len(r) else np.empty(0, np.int8)
            assert np.array_equal(got, oracle.decode_quals(data, r, add)), (trial, off)
    # a mirror that is not congruent to the buffer is refused, not silently skipped
    d = torch.frombuffer(bytearray(fqgen.fixed_records_np(100).tobytes()), dtype=torch.uint8).cuda()
    L = _lib.lib()
    ws = torch.empty(L.fqb_workspace_bytes(d.numel(), 0, 0) + 256, dtype=torch.uint8, device='cuda')
    own = torch.zeros(1, dtype=torch.int64, device='cuda')
    q = torch.empty(d.numel() + 32, dtype=torch.int8, device='cuda')
    bad = q[(d.data_ptr() - q.data_ptr()) % 16 + 1:]
    rc = L.fqb_shard_scan_decode(d.data_ptr(), d.numel(), d.numel(), 1, own.data_ptr(), None, 0, 0, bad.data_ptr(), -33,
                                 ws.data_ptr(), ws.numel(), 0, None)
    assert rc != 0


@pytest.mark.gpu
def test_sharded_parse_reports_halo_and_general(oracle):
    import torch
    import __graft_entry__ as g
    g.build()
    from fastqandfurious_b200 import _lib, shard
    data = fqgen.variable_records_np(40, 3, 'ont').tobytes()  # records of ~10 kb
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    rows, res = shard.parse_shards_local(d, [len(data) // 2], halo_bytes=64)
    assert rows is None and res.error == _lib.ERR_HALO
    rows, res = shard.parse_shards_local(d, [len(data) // 2], halo_bytes=len(data) // 2)
    assert rows is not None and np.array_equal(torch.cat(rows).cpu().numpy(), oracle.parse_chain(b'\n' + data, 0, -1)[0])
    data = fqgen.variable_records_np(400, 3, 'multiline').tobytes()
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    rows, res = shard.parse_shards_local(d, [len(data) // 2], halo_bytes=4096)
    assert rows is None and res.error == _lib.ERR_SHARD_GENERAL


def _stitch_general(rows, results):
    """Rows of all shards in order and the tail of the whole stream: the first shard (in order) whose result is
    not "continues in the next shard", else the last one's."""
    import torch
    table = torch.cat(rows).cpu().numpy() if rows else np.empty((0, 6), dtype=np.int64)
    for res in results:
        if res.error or res.tail_status != 6:
            return table, res
    return table, results[-1]


@pytest.mark.gpu
def test_sharded_general_path_matches_single_buffer(oracle):
    """Multi-line / damaged input cut at arbitrary bytes: the sharded general path (hand-over of the chain from
    shard to shard) emits exactly the single-buffer chain, with the same tail."""
    import torch
    import __graft_entry__ as g
    g.build()
    from fastqandfurious_b200 import shard
    rng = random.Random(17)
    cases = []
    for trial in range(14):
        kind = trial % 3
        if kind == 0:
            data = fqgen.variable_records_np(rng.randint(300, 900), trial, 'multiline').tobytes()
        elif kind == 1:
            data = fqgen.fastq_bytes(rng, rng.randint(300, 1200), read_len=(30, 120), header_len=(5, 30), wrap=rng.choice([0, 20, 60]),
                                     long_plus=0.3, trailing_newlines=rng.randint(0, 2), at_plus_bias=0.3)
        else:
            data = fqgen.mutate(rng, fqgen.fastq_bytes(rng, rng.randint(300, 900), read_len=(30, 120), header_len=(5, 30),
                                                        trailing_newlines=1, at_plus_bias=0.3), rng.randint(1, 4))
        cases.append(data)
    for trial, data in enumerate(cases):
        want, st, tail, resume = oracle.parse_chain(b'\n' + data, 0, -1)
        d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        world = rng.choice([2, 3, 4, 6])
        if len(data) < 3000 * world:
            world = 2
        cuts = sorted(rng.sample(range(1500, len(data) - 1500), world - 1))
        if min(b - a for a, b in zip([0] + cuts, cuts + [len(data)])) < 1200:
            continue
        rows, results = shard.parse_shards_local_general(d, cuts, halo_bytes=1100, epoch=trial + 1)
        table, last = _stitch_general(rows, results)
        assert not last.error, (trial, last.error)
        assert np.array_equal(table, want), (trial, world, cuts, len(table), len(want))
        assert last.tail_status == st, (trial, last.tail_status, st)
        # positions of the call that is not COMPLETE, as absolute stream offsets
        gi = results.index(last)
        goff = ([0] + cuts)[gi] - (1 if gi == 0 else 0)
        got_tail = [p + goff if p >= 0 else -1 for p in last.tail_pos]
        assert got_tail == [p - 1 if p >= 0 else -1 for p in tail.tolist()], (trial, gi)
        k0 = 0
        for r, res in zip(rows, results):
            assert res.reserved[0] == k0 or res.reserved[1] == 1
            k0 += len(r)


@pytest.mark.gpu
def test_single_shard_parser_fast_and_general_paths(oracle):
    """world == 1 (transport 'none'): ShardedParser.step and step_general on one shard equal the single-buffer parse
    (step_general used to fail before reaching the device: its hand-over epoch was only set up by the peer
    transports)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import __graft_entry__ as g
    g.build()
    import fastqandfurious_b200 as fq
    from fastqandfurious_b200 import shard
    for kind, force in (('illumina', False), ('multiline', True)):
        data = fqgen.variable_records_np(500, 21, kind).tobytes()
        want, st, tail, resume = oracle.parse_chain(b'\n' + data, 0, -1)
        plan = shard.ShardPlan(0, 1, [len(data)], 4096)
        sp = shard.ShardedParser(plan, 'cuda')
        assert sp.transport == 'none'
        sp.own().copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        table = torch.empty((len(want) + 8, 6), dtype=torch.int64, device='cuda')
        sp.step(table)
        if force:
            assert sp.needs_general()
            sp.step_general(table)
            sp.step_general(table)  # a second parse: fresh epoch, same answer
        res = sp.read()
        assert res.n_records == len(want) and res.tail_status == st
        assert np.array_equal(table[:res.n_records].cpu().numpy(), want)
        with pytest.raises(ValueError, match='int64'):
            sp.step(table.to(torch.int32))
        with pytest.raises(ValueError, match='int64'):
            sp.step_general(table[:, :5])


# ---- FASTA over byte-range shards ------------------------------------------------------------------------------------
def _oracle_fasta_parse(oracle):
    """The single-buffer FASTA call answered by the oracle, on CPU tensors (what device.parse_fasta_buffer does on the GPU)."""
    import torch

    def parse(buf, sentinel, goff):
        blob = (b'\n' if sentinel else b'') + buf.numpy().tobytes()
        table, st, tail, resume = oracle.fasta_chain(blob, 0, goff)
        return torch.from_numpy(table.reshape(-1, 4).copy()), st, tail.tolist(), resume
    return parse


def _fasta_stream(rng, n_records, header_only_run=0.0):
    out = []
    for k in range(n_records):
        if rng.random() < header_only_run:
            out.append(b'>h%d\n' % k * rng.randint(1, 9))  # a run of header-only records: consecutive "\n>" lines
            continue
        seq = bytes(rng.choice(b'ACGTN') for _ in range(rng.choice([0, 1, 30, 61, 200, 700])))
        out.append(b'>r%d some description\n' % k + fqgen._wrap(seq, rng.choice([0, 60])) + b'\n')
    return b''.join(out)[:-1 if rng.random() < 0.3 else None]


def test_fasta_shards_tile_the_whole_stream_chain(oracle):
    """The shards' rule on CPU (shard.parse_fasta_shards_local with the oracle as the single-buffer call): windows with
    a look-behind and a halo, ownership by the position of the "\\n>" newline, the run parity decided inside the
    look-behind -- the concatenated rows, the record count and the tail equal ONE chain over the whole stream, or the
    shards refuse (halo / look-behind too small); with generous windows they must answer."""
    import torch
    from fastqandfurious_b200 import shard
    rng = random.Random(41)
    parse = _oracle_fasta_parse(oracle)
    answered = refused = 0
    for trial in range(120):
        data = _fasta_stream(rng, rng.randint(5, 120), header_only_run=rng.choice([0.0, 0.0, 0.2]))
        if len(data) < 200:
            continue
        world = rng.choice([2, 3, 5])
        cuts = sorted(rng.sample(range(20, len(data) - 20), world - 1))
        generous = trial % 2 == 0
        halo, lb = (4000, 4000) if generous else (rng.choice([5, 60, 400]), rng.choice([1, 30, 400]))
        want, st, tail, resume = oracle.fasta_chain(b'\n' + data, 0, -1)
        t = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy())
        try:
            rows, n, status, tail_pos, res_off = shard.parse_fasta_shards_local(t, cuts, halo, lb, parse=parse)
        except shard.FastaShardError:
            refused += 1
            assert not generous or b'>h' in data, (data[:80], cuts)  # only long runs of header-only records may refuse
            continue
        answered += 1
        got = np.concatenate([r.numpy() for r in rows]) if n else np.empty((0, 4), dtype=np.int64)
        assert n == len(want) and np.array_equal(got, want), (data[:80], cuts, halo, lb)
        assert (status, tail_pos, res_off) == (st, tail.tolist(), resume), (data[:80], cuts, halo, lb)
    assert answered > 60 and refused > 3


def _gloo_fasta_worker(rank, world, port, data, own_lens, halo, lb, out):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle
    from fastqandfurious_b200 import shard
    sp = shard.ShardedFastaParser(rank, world, own_lens, halo, lb, parse=_oracle_fasta_parse(oracle))
    a = np.frombuffer(data, dtype=np.uint8)
    own = torch.from_numpy(a[sp.cuts[rank]:sp.cuts[rank + 1]].copy())
    k0, rows, n, status, tail_pos, resume = sp.parse(own)
    out.put((rank, k0, rows.numpy().tolist(), n, status, tail_pos, resume))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_fasta_parser_gloo(oracle, world):
    """ShardedFastaParser over gloo on CPU tensors (the oracle as the single-buffer call): the exchange of look-behind
    and halo bytes, the all-gather of the counts and the broadcast tail; every rank's rows start at its first record
    index and together they are the whole-stream chain."""
    import torch.multiprocessing as mp
    rng = random.Random(50 + world)
    data = _fasta_stream(rng, 300)
    cuts = [len(data) * (g + 1) // world + 11 * (g + 1) for g in range(world - 1)]
    own_lens = [b - a for a, b in zip([0] + cuts, cuts + [len(data)])]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_fasta_worker, args=(r, world, port, data, own_lens, 2000, 500, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    want, st, tail, resume = oracle.fasta_chain(b'\n' + data, 0, -1)
    got = []
    for rank, k0, rows, n, status, tail_pos, res_off in res:
        assert k0 == len(got) and n == len(want)
        assert (status, tail_pos, res_off) == (st, tail.tolist(), resume)
        got += rows
    assert got == want.tolist()


@pytest.mark.gpu
def test_fasta_shards_on_the_device(oracle):
    """The same rule with the CUDA parse as the single-buffer call: shards of one device-resident stream, one after the
    other (shard.parse_fasta_shards_local), against one oracle chain over the whole stream."""
    import torch
    from fastqandfurious_b200 import shard
    rng = random.Random(61)
    for trial in range(6):
        data = _fasta_stream(rng, 4000, header_only_run=0.0 if trial < 4 else 0.05)
        world = rng.choice([2, 3, 4, 7])
        cuts = sorted(rng.sample(range(1000, len(data) - 1000), world - 1))
        want, st, tail, resume = oracle.fasta_chain(b'\n' + data, 0, -1)
        d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
        rows, n, status, tail_pos, res_off = shard.parse_fasta_shards_local(d, cuts, 1 << 15, 1 << 13)
        got = np.concatenate([r.cpu().numpy() for r in rows])
        assert n == len(want) and np.array_equal(got, want), (trial, cuts)
        assert (status, tail_pos, res_off) == (st, tail.tolist(), resume), (trial, cuts)
    with pytest.raises(shard.FastaShardError):
        shard.parse_fasta_shards_local(d, [len(data) // 2], 16, 1 << 13)  # a halo shorter than a record
