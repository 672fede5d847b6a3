"""Time-bounded differential fuzzing of the CUDA paths against the oracle on a GPU box (tools/gpu_fuzz.sh).

The tests cover small damaged inputs and large clean ones; this covers what lies between: medium and large inputs
(up to > 4 Mi lines, so that every level of the general path's hierarchy is walked) with random damage anywhere,
through parse_buffer (automatic path choice and forced general path, with and without Phred decode), the sharded
protocols with local exchanges at random cuts, and FASTA.  A failing case is written to gpurun_out/fuzz_fail_<k>.bin
with its parameters next to it.  usage: FUZZ_SECONDS=60 FUZZ_SEED=1 python tests/fuzz_gpu.py"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))  # fqgen, algo_model

import numpy as np  # noqa: E402


def damage(rng, data, n_mut):
    """Random damage at random places: single bytes, ranges, whole lines dropped or doubled."""
    b = bytearray(data)
    for _ in range(n_mut):
        if len(b) < 4:
            break
        i = rng.randrange(len(b))
        op = rng.randrange(8)
        if op == 0:
            del b[i]
        elif op == 1:
            b.insert(i, rng.choice(b'\n\n@+AI'))
        elif op == 2:
            b[i] = rng.choice(b'\n@+AI\r')
        elif op == 3:
            del b[i:min(len(b), i + rng.randint(1, 400))]
        elif op == 4:  # drop the line that holds byte i
            s = b.rfind(b'\n', 0, i) + 1
            e = b.find(b'\n', i)
            e = len(b) if e < 0 else e + 1
            del b[s:e]
        elif op == 5:  # double it
            s = b.rfind(b'\n', 0, i) + 1
            e = b.find(b'\n', i)
            e = len(b) if e < 0 else e + 1
            b[s:s] = b[s:e]
        elif op == 6:
            b[i:i] = bytes(rng.choice(b'\n@+AI') for _ in range(rng.randint(1, 40)))
        else:
            b[i:i] = b'\n' * rng.randint(1, 3)
    if rng.random() < 0.2:
        del b[rng.randrange(max(1, len(b) - 2000), len(b) + 1):]
    return bytes(b)


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    import fastqandfurious_b200 as fq
    from fastqandfurious_b200 import shard
    import fqgen
    import oracle

    seconds = float(os.environ.get('FUZZ_SECONDS', '60'))
    seed = int(os.environ.get('FUZZ_SEED', '1'))
    rng = random.Random(seed)
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    bases = {k: fqgen.variable_records_np(n, seed + i, k).tobytes()
             for i, (k, n) in enumerate((('illumina', 6000), ('multiline', 6000), ('ont', 150)))}
    fasta_base = b''.join(b'>r%d some text\n' % k + fqgen._wrap(bytes(rng.choice(b'ACGTN') for _ in range(rng.choice([0, 30, 200, 700]))), rng.choice([0, 60, 70])) + b'\n'
                          for k in range(3000))
    counts, fails = {}, []

    def dev(data, offset=0):
        a = np.frombuffer(data, dtype=np.uint8)
        t = torch.full((len(a) + offset + 32,), 10, dtype=torch.uint8, device='cuda')
        if len(a):
            t[offset:offset + len(a)].copy_(torch.from_numpy(a.copy()))
        return t[offset:offset + len(a)]

    def fail(kind, data, params, why):
        k = len(fails)
        fails.append((kind, params, why))
        with open(os.path.join(out_dir, 'fuzz_fail_%d.bin' % k), 'wb') as fh:
            fh.write(data)
        with open(os.path.join(out_dir, 'fuzz_fail_%d.json' % k), 'w') as fh:
            json.dump({'kind': kind, 'params': params, 'why': why, 'seed': seed}, fh)
        print('FAIL', kind, params, why, flush=True)

    def case_parse():
        kind = rng.choice(['illumina', 'multiline', 'ont', 'multiline'])
        reps = rng.choice([1, 1, 2, 4, 12, 40 if kind == 'multiline' else 6])
        data = bases[kind] * reps
        n_mut = rng.choice([0, 1, 2, 5, 20, 100])
        data = damage(rng, data, n_mut)
        sentinel, goff, off = rng.choice([1, 1, 0]), rng.choice([-1, 0, 12345678901]), rng.randrange(16)
        force = rng.random() < 0.3
        decode = rng.random() < 0.3
        # the general path without its speculative single pass / with it as one CTA per chunk / as one warp per chunk
        spec = rng.choice([False, 'v1', True, True, True])
        params = dict(kind=kind, reps=reps, n_mut=n_mut, sentinel=sentinel, goff=goff, off=off, force=force, decode=decode,
                      spec=spec, n=len(data))
        blob = (b'\n' if sentinel else b'') + data
        want, st, tail, resume = oracle.parse_chain(blob, 0, goff)
        res = fq.parse_buffer(dev(data, off), sentinel=bool(sentinel), goff=goff, force_general=force, decode_quality=decode,
                              spec=spec)
        got = res.table.cpu().numpy()
        if res.n != len(want) or not np.array_equal(got, want):
            bad = int(np.argmax((got[:min(len(got), len(want))] != want[:min(len(got), len(want))]).any(1))) if len(got) and len(want) else -1
            return fail('parse', data, params, 'rows differ: %d vs %d, first bad %d' % (len(got), len(want), bad))
        if (res.tail_status, list(res.tail_pos), res.resume_offset) != (st, tail.tolist(), resume):
            return fail('parse', data, params, 'tail differs: %r vs %r' % ((res.tail_status, list(res.tail_pos), res.resume_offset), (st, tail.tolist(), resume)))
        if decode and len(got):
            q = res.qual.cpu().numpy()
            sel = got[rng.sample(range(len(got)), min(len(got), 300))]
            rel = sel - (goff + sentinel)  # rows are blob positions + goff; the mirror is indexed like the buffer
            gq = np.concatenate([q[a:b] for a, b in zip(rel[:, 4], rel[:, 5])])
            wq = oracle.decode_quals(data, rel)
            if not np.array_equal(gq, wq):
                return fail('parse', data, params, 'decoded qualities differ')
        key = (('parse_general_spec_v1' if spec == 'v1' else 'parse_general_spec') if res.spec else 'parse_general') if res.path == 2 else 'parse_fast'
        counts[key] = counts.get(key, 0) + 1

    def case_shard():
        kind = rng.choice(['illumina', 'multiline'])
        data = bases[kind] * rng.choice([1, 2, 3])
        general = kind == 'multiline' or rng.random() < 0.3
        if general:
            data = damage(rng, data, rng.choice([0, 1, 3, 10]))
        world = rng.choice([2, 3, 4, 8])
        halo = rng.choice([4096, 20000])
        lo = max(halo, 3000)
        if len(data) < (world + 1) * (lo + 2000):
            return
        cuts, prev = [], 0
        for k in range(world - 1):
            prev = rng.randrange(prev + lo, len(data) - (world - 1 - k) * lo - lo)
            cuts.append(prev)
        params = dict(kind=kind, general=general, world=world, cuts=cuts, halo=halo, n=len(data))
        want, st, tail, resume = oracle.parse_chain(b'\n' + data, 0, -1)
        d = dev(data)
        if general:
            rows, results = shard.parse_shards_local_general(d, cuts, halo_bytes=halo, epoch=rng.randint(1, 1000))
            last = next((r for r in results if r.error or r.tail_status != 6), results[-1])
            if last.error:
                if last.error == 5:  # FQB_ERR_HALO is a legitimate answer when damage made a record longer than the halo
                    counts['shard_general_halo'] = counts.get('shard_general_halo', 0) + 1
                    return
                return fail('shard_general', data, params, 'error %d' % last.error)
            got = torch.cat(rows).cpu().numpy()
            if not np.array_equal(got, want) or last.tail_status != st:
                return fail('shard_general', data, params, 'rows %d vs %d, status %d vs %d' % (len(got), len(want), last.tail_status, st))
            counts['shard_general'] = counts.get('shard_general', 0) + 1
        else:
            quals = [] if rng.random() < 0.5 else None
            rows, last = shard.parse_shards_local(d, cuts, halo_bytes=halo, fused=rng.random() < 0.5, epoch=rng.randint(1, 1000),
                                                  quals_out=quals, tail=rng.random() < 0.5)
            if rows is None:
                return fail('shard', data, params, 'error %d' % last.error)
            got = torch.cat(rows).cpu().numpy()
            if not np.array_equal(got, want) or last.tail_status != st:
                return fail('shard', data, params, 'rows %d vs %d' % (len(got), len(want)))
            if quals is not None:
                for r, (off, q) in zip(rows, quals):
                    r = r.cpu().numpy()[:200]
                    q = q.cpu().numpy()
                    if len(r) and not np.array_equal(np.concatenate([q[a - off:b - off] for a, b in zip(r[:, 4], r[:, 5])]),
                                                     oracle.decode_quals(data, r)):
                        return fail('shard', data, params, 'decoded qualities differ')
            counts['shard'] = counts.get('shard', 0) + 1

    def case_fasta():
        data = fasta_base * rng.choice([1, 2, 8])
        if rng.random() < 0.5:
            i = rng.randrange(len(data))
            data = data[:i] + b'>\n' * rng.choice([1, 2, 33, 5000, 70001]) + data[i:]
        data = damage(rng, data, rng.choice([0, 1, 5, 50]))
        sentinel, goff, off = rng.choice([1, 0]), rng.choice([-1, 0, 777]), rng.randrange(16)
        params = dict(sentinel=sentinel, goff=goff, off=off, n=len(data))
        want, st, tail, resume = oracle.fasta_chain((b'\n' if sentinel else b'') + data, 0, goff)
        res = fq.parse_fasta_buffer(dev(data, off), sentinel=bool(sentinel), goff=goff)
        if res.n != len(want) or not np.array_equal(res.table.cpu().numpy(), want):
            return fail('fasta', data, params, 'rows differ: %d vs %d' % (res.n, len(want)))
        if (res.tail_status, res.tail_pos, res.resume_offset) != (st, tail.tolist(), resume):
            return fail('fasta', data, params, 'tail differs')
        counts['fasta'] = counts.get('fasta', 0) + 1

    host_parsers = {}

    def case_stream():
        """The chunked host front-ends (refill and end-of-stream rules applied per chunk, tails carried over) against
        the oracle's readfastq_iter: readfastq_table on a file object and HostParser.parse on a host tensor."""
        import io
        kind = rng.choice(['illumina', 'multiline', 'ont'])
        data = bases[kind]
        if rng.random() < 0.5:
            data = data[:rng.randrange(1, len(data))]
        data = damage(rng, data, rng.choice([0, 0, 1, 2, 5]))
        want, err, err_byte = oracle.readfastq(data)
        msg = {0: None, 1: 'Incomplete final quality string at byte', 2: 'Incomplete entry at byte %i' % err_byte,
               3: 'Entry is invalid at byte %i' % err_byte}[err]
        which = rng.choice(['table', 'host'])
        chunk = rng.choice([1 << 26, 1 << 20, 300_000, 70_001]) if kind != 'ont' else rng.choice([1 << 26, 1 << 21])
        params = dict(kind=kind, which=which, chunk=chunk, n=len(data), err=err)
        try:
            if which == 'table':
                rows = fq.readfastq_table(io.BytesIO(data), rng.choice([1, 100, 65536]), device_chunk=chunk)
            else:
                hp = host_parsers.get(chunk)
                if hp is None:
                    hp = host_parsers[chunk] = fq.device.HostParser('cuda', chunk_bytes=chunk)
                rows = hp.parse(torch.frombuffer(bytearray(data), dtype=torch.uint8)).copy()
            got_msg = None
        except ValueError as e:
            rows, got_msg = e.rows, str(e)
        if len(rows) != len(want) or not np.array_equal(rows, want):
            return fail('stream', data, params, 'rows differ: %d vs %d' % (len(rows), len(want)))
        if got_msg != msg:
            return fail('stream', data, params, 'error differs: %r vs %r' % (got_msg, msg))
        counts['stream_' + which] = counts.get('stream_' + which, 0) + 1

    t0 = time.time()
    n = 0
    only = os.environ.get('FUZZ_ONLY')
    menu = {'parse': [case_parse], 'shard': [case_shard], 'fasta': [case_fasta], 'stream': [case_stream]}.get(
        only, [case_parse, case_parse, case_shard, case_fasta, case_stream])
    while time.time() - t0 < seconds and len(fails) < 5:
        rng.choice(menu)()
        n += 1
    print('fuzz: %d cases in %.0f s, seed %d, passed by kind %s, failures %d' % (n, time.time() - t0, seed, json.dumps(counts, sort_keys=True), len(fails)))
    sys.exit(1 if fails else 0)


if __name__ == '__main__':
    main()
