"""The synthetic streams of BASELINE.json configs[2..4] (csrc/fq_synth.cuh): the library's host-side record
generator and the numpy twin agree byte for byte, the twin's truth table is what the oracle parses (CPU); on the GPU the
device generator equals the twin, and the parser -- single buffer and sharded at arbitrary bytes, long reads
included -- reproduces the generator's truth."""
import random

import numpy as np
import pytest

import fqgen


@pytest.fixture(scope='module')
def synth():
    import __graft_entry__ as g
    g.build()
    from fastqandfurious_b200 import synth as m
    return m


@pytest.mark.parametrize('kind,n', [('illumina', 2500), ('ont', 60), ('multiline', 400)])
def test_host_record_matches_numpy_twin_and_oracle(oracle, synth, kind, n):
    qt = synth.ont_qtable()
    assert len(qt) == 4097 and qt[0] == 200 and qt[-1] == 500000 and (np.diff(qt) >= 0).all()
    data, rows, offs = fqgen.synth_records_np(kind, n, qtable=qt)
    for k in range(n):
        b, meta = synth.host_record(kind, k, int(offs[k]))
        assert b == data[offs[k]:offs[k + 1]].tobytes(), (kind, k)
        hl, rl, sb, pl = meta
        assert (hl, sb) == (rows[k, 1] - rows[k, 0], rows[k, 3] - rows[k, 2])
        assert pl == rows[k, 4] - rows[k, 3] - 2 and rows[k, 5] - rows[k, 4] == sb
    want, err, _ = oracle.readfastq(data.tobytes())
    assert err == 0 and np.array_equal(want, rows)
    if kind == 'illumina':  # variable-width headers, '@' / '+' hazards at line starts
        widths = {int(w) for w in rows[:, 1] - rows[:, 0]}
        assert len(widths) >= 3
        firstq = data[rows[:, 4]]
        assert (firstq == ord('@')).any() or (firstq == ord('+')).any()
    if kind == 'multiline':  # wrapped like data/test_multiline.fq, long '+' lines like test_longqualityheader.fq
        assert ((rows[:, 4] - rows[:, 3]) > 3).any() and ((rows[:, 4] - rows[:, 3]) == 3).any()
        assert (rows[:, 3] - rows[:, 2] > 150).all()


def test_ont_lengths_follow_the_quantile_table(synth):
    qt = synth.ont_qtable()
    seed = fqgen.SYNTH_SEEDS['ont']
    lens = np.array([fqgen.synth_record_meta('ont', k, seed, qt)[1] for k in range(30000)])
    assert 9500 < lens.mean() < 10600 and lens.min() >= 200 and lens.max() <= 500000
    assert (lens > 30000).sum() > 50  # the long tail is there: rows that span many 16 KiB tiles


# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def fq():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import __graft_entry__ as g
    g.build()
    import fastqandfurious_b200 as m
    return m


@pytest.mark.gpu
@pytest.mark.parametrize('kind,n', [('illumina', 3000), ('ont', 150), ('multiline', 3000)])
def test_device_generator_equals_twin_and_parser_equals_truth(fq, synth, oracle, kind, n):
    import torch
    data, rows, offs = fqgen.synth_records_np(kind, n, qtable=synth.ont_qtable())
    st = synth.SynthStream(kind, n)
    assert st.total == len(data) and np.array_equal(st.off.cpu().numpy(), offs)
    buf = st.fill()
    assert np.array_equal(buf.cpu().numpy(), data)
    rng = random.Random(2)
    for _ in range(12):  # windows at arbitrary bytes (one shard per GPU generates its own)
        a = rng.randrange(len(data))
        b = min(len(data), a + rng.choice([1, 15, 16, 17, 511, 513, 70000]))
        off = rng.randrange(16)
        out = torch.zeros(b - a + off + 16, dtype=torch.uint8, device='cuda')
        st.fill(a, b - a, out=out[off:off + b - a])
        assert np.array_equal(out[off:off + b - a].cpu().numpy(), data[a:b])
        assert int(out[:off].sum()) == 0 and int(out[off + b - a:].sum()) == 0
    truth = st.truth(0, n)
    assert np.array_equal(truth.cpu().numpy(), rows)
    res = fq.parse_buffer(buf, cap=n + 8)
    # the last record ends with the buffer: status 5 like the reference's entrypos, completed by the end-of-stream rule
    assert res.n == n - 1 and res.path == (2 if kind == 'multiline' else 1) and res.tail_status == fq.MISSING_QUAL_END
    assert [p - 1 for p in res.tail_pos[:5]] == rows[-1, :5].tolist()
    assert st.mismatches(res.table, 0, chunk=1000) == 0
    import io
    assert np.array_equal(fq.readfastq_table(io.BytesIO(data.tobytes())), rows)
    from fastqandfurious_b200 import shard
    job = shard.SynthJob(kind, len(data) + 5, 0, 1, 'cuda')
    job.step()
    assert job.stream.n == n and job.verify_local() == n
    wins = job.seam_windows(n_seams=3, half=60000)
    assert sum(len(w[3]) for w in wins) >= 3
    for k_a, k_b, wdata, wtruth in wins:
        got, err, _ = oracle.readfastq(wdata)
        assert err == 0 and np.array_equal(got, wtruth)
    # ownership bookkeeping used by the sharded benches
    cut = int(offs[n // 2]) + 5
    assert st.records_from(0, cut) == (0, n // 2 + 1) and st.records_from(cut, len(data)) == (n // 2 + 1, n)
    assert st.records_from(int(offs[7]), int(offs[9])) == (8, 10)  # pos0 - 1 in [off7, off9): records 8, 9
    s2 = synth.SynthStream.for_bytes(kind, len(data) - 3)
    assert s2.n == n - 1 and s2.total == int(offs[n - 1])


@pytest.mark.gpu
def test_sharded_parse_of_reads_beyond_100kb(fq, synth, oracle):
    """BASELINE config 4 at its stressing size (src/fastqandfurious.py:219-232: a buffer must hold the largest entry;
    here: the halo): reads of 100-450 kb, the stream cut at arbitrary bytes into shards with a 1 MiB halo -- plain and
    fused exchange -- equal the generator's truth and the oracle's chain; a halo shorter than a record is reported."""
    import torch
    from fastqandfurious_b200 import shard, _lib
    qt = np.linspace(100000, 450000, 4097).astype(np.int32)
    st = synth.SynthStream.for_bytes('ont', 24 << 20, seed=77, qtable=qt)
    buf = st.fill()
    n = st.n
    truth = st.truth(0, n).cpu().numpy()
    assert (truth[:, 3] - truth[:, 2]).min() >= 100000
    host = buf.cpu().numpy().tobytes()
    want, err, _ = oracle.readfastq(host)
    assert err == 0 and np.array_equal(want, truth)
    rng = random.Random(9)
    halo = 1 << 20
    for trial in range(4):
        world = rng.choice([2, 3, 4])
        while True:
            cuts = sorted(rng.sample(range(halo, st.total - 1), world - 1))
            lens = [b - a for a, b in zip([0] + cuts, cuts + [st.total])]
            if min(lens[:-1]) >= halo and min(lens[1:-1] or [halo]) >= halo:
                break
        for fused in (False, True):
            rows, last = shard.parse_shards_local(buf, cuts, halo, fused=fused, epoch=trial + 1)
            assert rows is not None, (trial, cuts, last.error)
            got = torch.cat(rows).cpu().numpy()
            # (the stream's last record ends with the buffer: MISSING_QUAL_END, completed by the end-of-stream rule)
            assert np.array_equal(got, truth[:-1]), (trial, cuts, fused)
            assert last.tail_status == 5 and [p + cuts[-1] for p in last.tail_pos[:5]] == truth[-1, :5].tolist()
            bounds = [0] + cuts + [st.total]
            for g, r in enumerate(rows):  # every shard emitted exactly the records it owns
                k_lo, k_hi = st.records_from(bounds[g], bounds[g + 1])
                assert len(r) == k_hi - k_lo - (1 if g == world - 1 else 0), (trial, g)
    rows, last = shard.parse_shards_local(buf, [st.total // 2], 64 << 10)
    assert rows is None and last.error == _lib.ERR_HALO


@pytest.mark.gpu
def test_speculative_pass_answers_clean_multiline_at_scale(fq, synth):
    """Config 5's stream at 512 MiB (about 1.1 M records, 5 000 chunks, 170 000 walker regions): the speculative pass
    must ANSWER -- a run-up too short for the walkers lets a few of them start on a quality line that begins with '@'
    and the whole buffer falls back to the exact path (same rows, twice the time); and the long-read mode of the
    emit kernel must give the truth on ONT-like input of the same size."""
    import torch
    from fastqandfurious_b200 import device, shard
    job = shard.SynthJob('multiline', 1 << 29, 0, 1, 'cuda')
    job.prepare()
    assert not job.exact, 'the speculative general pass declined clean wrapped records'
    job.step()
    torch.cuda.synchronize()
    res = device.read_result(job.result)
    assert res.path == 2 and res.reserved[1] == 1
    assert job.verify_local() == job.stream.n
    job.free()
    job = shard.SynthJob('ont', 1 << 29, 0, 1, 'cuda')
    job.step()
    assert job.verify_local() == job.stream.n
    job.free()
