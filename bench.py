#!/usr/bin/env python
"""bench.py -- FASTQ GB/s (and Mrecords/s) of the FASTQ-buffer -> offset-table hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over the whole synthetic buffer resident on each GPU.  At N=1 the
workload is BASELINE.json configs[1]: 1 GiB of fixed-length 150 bp single-line FASTQ (337 B/record,
3 186 177 records), generated on the device (outside the timed region).  With N>1 (torchrun, one
process per GPU) every rank holds one such shard of a single logical N-GiB stream cut at arbitrary
byte positions, parses it, and the shards are stitched with one neighbour exchange (weak scaling).

One JSON line on stdout (rank 0):
  value       whole-job GB/s with the input resident in HBM (CUDA events, max over ranks)
  e2e         same metric through the host-buffer API (pinned host memory -> device -> offset table
              back on the host; H2D / D2H inside the timed region)
  roofline    the scan kernel against the measured HBM peak (algorithmic bytes / its event-timed
              duration), cpu_baseline: the reference's C extension on the host cores (bounded sample)
`--impl reference` times the unmodified reference (oracle/_ref: its C `entrypos` + `readfastq_iter` +
`entryfunc_abspos`) on all host cores on a bounded sample of the same workload.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'fastq-and-furious_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

REC_BYTES = 337  # '@SIM:%012d 1:N:0:ACGTACGT' (32) \n 150 \n + \n 150 \n
WORKLOADS = {
    # name: (description, bytes per GPU)
    'fixed150_1g': ('synthetic 1 GiB single-line FASTQ, 150 bp fixed-length reads (BASELINE.json configs[1])', 1 << 30),
    'fixed150_64m': ('64 MiB of the same shape (quick check)', 1 << 26),
}
METRIC = 'FASTQ GB/s (input bytes parsed to the per-record offset table per second)'


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [s for s in self.samples if self.t0 is None or self.t0 - 0.05 <= s[0] <= self.t1 + 0.05] or self.samples
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for _, line in rows:
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref = the unmodified reference; else the oracle port)
# ---------------------------------------------------------------------------------------------------
_CPU_SAMPLE = None  # bytes-like, set before the worker pool is forked (workers slice it; nothing is pickled)


def _cpu_worker(args):
    lo, hi, repeats, use_ref, fbufsize = args
    import io
    from array import array
    import oracle
    data = bytes(memoryview(_CPU_SAMPLE)[lo:hi])
    nrec = 0
    t0 = time.perf_counter()
    if use_ref:
        mod, cext = oracle.reference()
        for _ in range(repeats):
            out = array('q')
            for pos in mod.readfastq_iter(io.BytesIO(data), fbufsize, entryfunc=mod.entryfunc_abspos,
                                          entrypos=cext.entrypos):
                out.extend(pos)
            nrec += len(out) // 6
    else:
        for _ in range(repeats):
            table, err, _ = oracle.readfastq(data)
            nrec += len(table)
    return time.perf_counter() - t0, nrec, len(data) * repeats


def cpu_reference_rate(sample, procs, repeats, starts=None, fbufsize=2 ** 16):
    """Whole-file GB/s of the reference's own path (readfastq_iter + C entrypos + entryfunc_abspos,
    fbufsize 2**16 as in src/demo/benchmark.py:26-27) over `sample` (bytes), split into `procs`
    record-aligned slices, one process each (the reference itself is single-threaded).  `starts`: sorted
    byte offsets of the record starts (+ the end) for variable geometry; default: REC_BYTES records."""
    global _CPU_SAMPLE
    import multiprocessing as mp
    import oracle
    oracle.build()
    use_ref = oracle.reference() is not None
    if starts is None:
        nrec_total = len(sample) // REC_BYTES
        per = max(1, nrec_total // procs)
        bounds = [min(i * per, nrec_total) * REC_BYTES for i in range(procs)] + [min(procs * per, nrec_total) * REC_BYTES]
    else:
        nrec_total = len(starts) - 1
        per = max(1, nrec_total // procs)
        bounds = [int(starts[min(i * per, nrec_total)]) for i in range(procs)] + [int(starts[min(procs * per, nrec_total)])]
    jobs = [(bounds[i], bounds[i + 1], repeats, use_ref, fbufsize) for i in range(procs) if bounds[i + 1] > bounds[i]]
    _CPU_SAMPLE = sample
    ctx = mp.get_context('fork')
    t0 = time.perf_counter()
    if len(jobs) == 1:
        results = [_cpu_worker(jobs[0])]
    else:
        with ctx.Pool(len(jobs)) as pool:
            results = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    _CPU_SAMPLE = None
    nbytes = sum(r[2] for r in results)
    nrec = sum(r[1] for r in results)
    worker_wall = max(r[0] for r in results)
    return {'gbs': nbytes / worker_wall / 1e9, 'mrec_s': nrec / worker_wall / 1e6, 'bytes': nbytes, 'wall_s': wall,
            'cores': len(jobs), 'kind': 'reference' if use_ref else 'port'}


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return None


def host_sample(nbytes):
    """First `nbytes` of the workload on the host (numpy twin of the device generator)."""
    import fqgen
    return fqgen.fixed_records_np(nbytes // REC_BYTES).tobytes()


def cpu_single_process_legs(sample, seconds=2.5):
    """The single-process legs of the reference's own benchmark on `sample` (kind 'reference' only):
    (i) readfastq_iter + C entrypos + entryfunc_abspos at fbufsize 2**16 (src/demo/benchmark.py:26-27) and 50 000
    (:415-419); (ii) a bare C entrypos loop over the whole buffer (upper bound of the C path); (iv) the per-record
    Phred decode recipe array('b').frombytes + arrayadd_b(-33) (:159-163); plus the reference's own unit,
    sequence letters MB/s (:16-23, doc/performance.rst:24-25).  Each leg parses a prefix of `sample` sized to take
    about `seconds`."""
    import io as _io
    from array import array
    import oracle
    ref = oracle.reference()
    if ref is None:
        return None
    mod, cext = ref
    out = {}
    sample = bytes(sample)

    def timed(fn, data):
        t0 = time.perf_counter()
        nrec, nseq = fn(data)
        dt = time.perf_counter() - t0
        return {'gbs': len(data) / dt / 1e9, 'mrec_s': nrec / dt / 1e6, 'seq_letters_mb_s': nseq / dt / 1e6,
                'bytes': len(data), 'seconds': dt}

    def sized(fn):
        probe = sample[:REC_BYTES * 20000]
        r = timed(fn, probe)
        want = int(min(len(sample), max(len(probe), r['gbs'] * 1e9 * seconds)))
        return timed(fn, sample[:want // REC_BYTES * REC_BYTES])

    def iter_abspos(fbufsize):
        def fn(data):
            n = seq = 0
            for pos in mod.readfastq_iter(_io.BytesIO(data), fbufsize, entryfunc=mod.entryfunc_abspos,
                                          entrypos=cext.entrypos):
                n += 1
                seq += pos[3] - pos[2]
            return n, seq
        return fn

    def bare(data):
        buf = b'\n' + data
        pos = array('q', [-1] * 6)
        off = n = seq = 0
        ep = cext.entrypos
        while ep(buf, off, pos) == 6:
            off = pos[5] - 1
            n += 1
            seq += pos[3] - pos[2]
        return n, seq

    def decode(data):
        n = seq = 0
        for h, sq, q in mod.readfastq_iter(_io.BytesIO(data), 2 ** 16, entryfunc=mod.entryfunc, entrypos=cext.entrypos):
            a = array('b')
            a.frombytes(q)
            cext.arrayadd_b(a, -33)
            n += 1
            seq += len(sq)
        return n, seq

    out['readfastq_iter_abspos_fbufsize_65536'] = sized(iter_abspos(2 ** 16))
    out['readfastq_iter_abspos_fbufsize_50000'] = sized(iter_abspos(50000))
    out['bare_entrypos_loop'] = sized(bare)
    out['readfastq_iter_entryfunc_plus_arrayadd_b_decode'] = sized(decode)
    return out


def dropin_rate(fq, data, seconds=2.5):
    """Records/s of OUR drop-in generator called exactly like the reference's:
    readfastq_iter(io.BytesIO(data), 2**16, entryfunc=entryfunc_abspos), one Python object per record."""
    import io as _io

    def run(d):
        t0 = time.perf_counter()
        n = seq = 0
        for pos in fq.readfastq_iter(_io.BytesIO(d), 2 ** 16, entryfunc=fq.entryfunc_abspos):
            n += 1
            seq += pos[3] - pos[2]
        dt = time.perf_counter() - t0
        return {'gbs': len(d) / dt / 1e9, 'mrec_s': n / dt / 1e6, 'seq_letters_mb_s': seq / dt / 1e6, 'bytes': len(d),
                'seconds': dt, 'records': n}
    probe = bytes(data[:REC_BYTES * 100000])
    run(probe)
    r = run(probe)
    want = int(min(len(data), max(len(probe), r['gbs'] * 1e9 * seconds)))
    r = run(bytes(data[:want // REC_BYTES * REC_BYTES]))
    r['api'] = 'fastqandfurious_b200.readfastq_iter(io.BytesIO(data), 2**16, entryfunc=entryfunc_abspos)'
    return r


def cpu_baseline_block(sample_bytes, total_bytes):
    procs = os.cpu_count() or 1
    sample = host_sample(sample_bytes)
    repeats = max(1, int(total_bytes // max(1, len(sample))))
    r = cpu_reference_rate(sample, procs, repeats)
    return r, ('first %d MiB of the workload, %d record-aligned slices x %d passes = %.1f GiB parsed, '
               'readfastq_iter(fbufsize=65536)+C entrypos+entryfunc_abspos' %
               (sample_bytes >> 20, r['cores'], repeats, r['bytes'] / 2 ** 30)), sample


def static_config(workload):
    """The `config` object both arms print: the workload BASELINE.json names, nothing run dependent."""
    desc, nbytes = WORKLOADS[workload]
    return {'workload': workload, 'description': desc, 'bytes_per_gpu': nbytes // REC_BYTES * REC_BYTES,
            'records_per_gpu': nbytes // REC_BYTES, 'record_bytes': REC_BYTES,
            'l2': 'input (1 GiB/GPU) is larger than L2 (126 MB); no flush needed'}


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    desc, nbytes = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    sample = host_sample(min(nbytes, args.ref_sample))  # the first 1 GiB of the workload
    # each step parses the sample once, one record-aligned slice per core; bounded so K+W steps end within minutes
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(sample, procs, 1)
        if i >= args.warmup:
            times.append((r['bytes'], r['bytes'] / r['gbs'] / 1e9, r['mrec_s']))
    tot_b = sum(t[0] for t in times)
    tot_s = sum(t[1] for t in times)
    gbs = tot_b / tot_s / 1e9
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': gbs, 'unit': 'GB/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * tot_s / max(1, len(times)), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'mrecords_per_s': sum(t[2] for t in times) / max(1, len(times)),
        'config': static_config(args.workload),
        'cpu_baseline': {'value': gbs, 'unit': 'GB/s', 'cores': r['cores'], 'kind': r['kind'], 'cpu_model': cpu_model(),
                         'sample': 'each step: first %d MiB of the workload in %d record-aligned slices, one process '
                                   'per core, readfastq_iter(fbufsize=65536)+C entrypos+entryfunc_abspos' %
                                   (len(sample) >> 20, r['cores'])},
        'e2e': {'value': gbs, 'unit': 'GB/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def _time_steps(torch, fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def measure_extras(fq, device, _lib, torch, buf, table, args):
    """GB/s of the other device paths on 1 GiB inputs: fused-call Phred decode, ONT-like long reads (fast
    path), wrapped multi-line records with long '+' headers (general path).  Inputs for the last two are
    64 MiB record-aligned numpy blocks (tests/fqgen.py) repeated on the device."""
    import fqgen
    out = {}
    steps = max(5, min(args.steps, 30))
    dev = buf.device
    result = torch.empty(16, dtype=torch.int64, device=dev)
    flags = _lib.FLAG_CFG(args.cfg) | _lib.FLAG_FAST_ONLY
    qual = torch.empty(buf.numel(), dtype=torch.int8, device=dev)
    ms = _time_steps(torch, lambda: device.parse_raw(buf, 1, -1, table, qual, -33, result, flags), steps)
    out['fixed150_1g_with_phred_decode'] = {'gbs': buf.numel() / ms / 1e6, 'ms_per_step': ms, 'path': 'fast4'}
    del qual
    # consumers of the offset table: index replay of every sequence / decoded quality, Phred sums
    from fastqandfurious_b200 import consume
    res = fq.parse_buffer(buf, cap=table.shape[0], table=table)
    rows = res.table
    lens = consume.field_lengths(rows, 'sequence')
    offsets = consume.exclusive_scan(lens)
    total = int(offsets[-1].item())
    packed = torch.empty(total, dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    sums = torch.empty(rows.shape[0], dtype=torch.int64, device=dev)
    L = _lib.lib()

    def gather(field, add):
        _lib.check(L.fqb_gather_fields(buf.data_ptr(), buf.numel(), 0, rows.data_ptr(), rows.shape[0], None, rows.shape[0],
                                       field, offsets.data_ptr(), packed.data_ptr(), add & 0xff, status.data_ptr(),
                                       device._stream()), 'fqb_gather_fields')

    for name, fn, moved in (('index_replay_sequences_1g', lambda: gather(1, 0), 2 * total + 16 * rows.shape[0]),
                            ('index_replay_phred_decoded_1g', lambda: gather(2, -33), 2 * total + 16 * rows.shape[0]),
                            ('phred_sums_1g', lambda: _lib.check(L.fqb_field_sums(
                                buf.data_ptr(), buf.numel(), 0, rows.data_ptr(), rows.shape[0], None, rows.shape[0], 2,
                                (-33) & 0xff, sums.data_ptr(), status.data_ptr(), device._stream()), 'fqb_field_sums'),
                             total + 24 * rows.shape[0])):
        ms = _time_steps(torch, fn, steps)
        out[name] = {'ms_per_step': ms, 'field_bytes': total, 'records': int(rows.shape[0]),
                     'gbs_of_field_bytes': total / ms / 1e6, 'hbm_gbs': moved / ms / 1e6}
    # 2-bit packed sequences of all records (fqb_pack_2bit)
    try:
        slot_off = consume.exclusive_scan(((lens + 15) // 16) * 4)
        ptotal = int(slot_off[-1].item())
        pk = torch.empty(ptotal, dtype=torch.uint8, device=dev)
        nb = torch.empty(rows.shape[0], dtype=torch.int64, device=dev)
        ms = _time_steps(torch, lambda: _lib.check(L.fqb_pack_2bit(
            buf.data_ptr(), buf.numel(), 0, rows.data_ptr(), rows.shape[0], None, rows.shape[0], slot_off.data_ptr(),
            pk.data_ptr(), nb.data_ptr(), None, status.data_ptr(), device._stream()), 'fqb_pack_2bit'), steps)
        out['pack_2bit_sequences_1g'] = {'ms_per_step': ms, 'field_bytes': total, 'packed_bytes': ptotal,
                                         'records': int(rows.shape[0]), 'gbs_of_field_bytes': total / ms / 1e6}
        del pk, nb, slot_off
    except Exception as exc:
        out['pack_2bit_sequences_1g'] = {'failed': repr(exc)}
    assert int(status.item()) == 0
    del packed, lens, offsets, sums, status, rows, res
    for name, kind, nrec in (('ont10k_1g', 'ont', 6000), ('multiline_1g', 'multiline', 120000)):
        base = fqgen.variable_records_np(nrec, 31, kind)
        reps = max(1, (1 << 30) // len(base))
        d = torch.from_numpy(base).to(dev).repeat(reps)
        res = fq.parse_buffer(d, cap=nrec * reps + 64)
        tab = res.table_full
        if res.path == 1:
            ms = _time_steps(torch, lambda: device.parse_raw(d, 1, -1, tab, None, 0, result, flags), steps)
        else:
            gflags = _lib.FLAG_CFG(args.cfg) | _lib.FLAG_FORCE_GENERAL | (_lib.FLAG_SPEC_ONLY if res.spec else _lib.FLAG_NO_SPEC)
            ml = 0 if res.spec else res.n_lines + 64
            ms = _time_steps(torch, lambda: device.parse_raw(d, 1, -1, tab, None, 0, result, gflags, max_lines=ml), steps)
        out[name] = {'gbs': d.numel() / ms / 1e6, 'ms_per_step': ms,
                     'path': 'fast4' if res.path == 1 else ('general (speculative pass)' if res.spec else 'general (exact)'),
                     'records': int(res.n), 'bytes': int(d.numel())}
        del d, tab, res
    # the drop-in call on a plain Python file object: readfastq_table(io.BytesIO(...)), 1 GiB
    import io as _io
    host_bytes = buf.cpu().numpy().tobytes()
    fq.readfastq_table(_io.BytesIO(host_bytes[:REC_BYTES * 200000]))  # warm-up (record aligned)
    t0 = time.perf_counter()
    tab_h = fq.readfastq_table(_io.BytesIO(host_bytes))
    dt = time.perf_counter() - t0
    out['readfastq_table_bytesio_1g'] = {'gbs': len(host_bytes) / dt / 1e9, 'seconds': dt, 'records': int(len(tab_h)),
                                         'api': 'fastqandfurious_b200.readfastq_table(io.BytesIO(data)) -> int64[n,6] ndarray'}
    # ... and on a regular file (tmpfs when there is one: the page cache, not a disk), read by several threads
    try:
        import tempfile
        tdir = '/dev/shm' if os.path.isdir('/dev/shm') else None
        with tempfile.NamedTemporaryFile(dir=tdir, delete=False) as tf:
            tf.write(host_bytes)
            tpath = tf.name
        try:
            best = None
            for _ in range(2):
                with open(tpath, 'rb') as fh:
                    t0 = time.perf_counter()
                    tab_f = fq.readfastq_table(fh)
                    dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out['readfastq_table_file_1g'] = {'gbs': len(host_bytes) / best / 1e9, 'seconds': best, 'records': int(len(tab_f)),
                                              'api': "fastqandfurious_b200.readfastq_table(open(path, 'rb')) -> int64[n,6] ndarray"}
        finally:
            os.unlink(tpath)
    except OSError as exc:
        out['readfastq_table_file_1g'] = {'skipped': repr(exc)}
    del host_bytes, tab_h
    # FASTA: records of 300 bases wrapped at 60 columns, 1 GiB
    import numpy as np
    rng = np.random.default_rng(6)
    nrec = 190000
    rec = bytearray()
    seqs = rng.choice(np.frombuffer(b'ACGT', dtype=np.uint8), size=(nrec, 5, 60))
    for k in range(nrec):
        rec += b'>read%07d sample\n' % k
        rec += b'\n'.join(bytes(row) for row in seqs[k]) + b'\n'
    base = np.frombuffer(bytes(rec), dtype=np.uint8)
    reps = max(1, (1 << 30) // len(base))
    d = torch.from_numpy(base.copy()).to(dev).repeat(reps)
    res = device.parse_fasta_buffer(d)
    ml, cap4 = int(res.n_lines) + 64, int(res.n) + 64
    tab4 = torch.empty((cap4, 4), dtype=torch.int64, device=dev)
    ws = torch.empty(L.fqb_fasta_workspace_bytes(d.numel(), ml, 0) + 256, dtype=torch.uint8, device=dev)
    ms = _time_steps(torch, lambda: _lib.check(L.fqb_parse_fasta(d.data_ptr(), d.numel(), 1, -1, tab4.data_ptr(), cap4,
                                                                 result.data_ptr(), ws.data_ptr(), ws.numel(), ml, 0,
                                                                 device._stream()), 'fqb_parse_fasta'), steps)
    out['fasta_1g'] = {'gbs': d.numel() / ms / 1e6, 'ms_per_step': ms, 'records': int(res.n), 'bytes': int(d.numel()),
                       'lines': int(res.n_lines)}
    del d, tab4, ws, res
    device._ws_cache.clear()
    torch.cuda.empty_cache()
    return out


CONFIGS = {
    # extras key: (BASELINE.json config, synthetic kind, default GiB of the WHOLE stream, halo bytes at N > 1)
    'cfg3_illumina150_64g': ('configs[2]: 64 GiB Illumina-like 150 bp, variable-width headers, chunk-sharded over N GPUs',
                             'illumina', 64.0, 1 << 20),
    'cfg4_ont10k_8g': ('configs[3]: ONT-like long reads, 10 kb mean (Gamma k=2), 200 b - 500 kb', 'ont', 8.0, 4 << 20),
    'cfg5_multiline_8g': ("configs[4]: reads wrapped at 60 columns + long '+' lines, 8 GiB (general path)", 'multiline', 8.0,
                          1 << 20),
}


def check_seam_windows(job):
    """The generator's truth in windows around the shard seams against the compiled reference (oracle/_ref;
    the oracle port when it is absent): readfastq_iter + C entrypos + entryfunc_abspos over the window's bytes.
    Outside every timed region; the checker, not the product.  Returns (windows, records) checked on this rank."""
    import numpy as np
    import oracle
    wins = recs = 0
    for k_a, k_b, data, truth in job.seam_windows():
        if oracle.reference() is not None:
            got = oracle.reference_abspos(data, fbufsize=max(2 ** 16, 4 << 20))
        else:
            got, err, _ = oracle.readfastq(data)
        if not np.array_equal(got, truth):
            raise AssertionError('%s: the reference parses records %d..%d around a seam differently from the generator truth'
                                 % (job.kind, k_a, k_b))
        wins += 1
        recs += len(truth)
    return wins, recs


def measure_config(key, args, rank, world, dev, torch, dist, peak):
    """One BASELINE config as an extras entry: STRONG scaling (the stream is fixed, each rank parses total / N bytes
    + halo), every row of every rank verified against the generator truth, seam windows against the reference.
    Called on every rank; after each phase the ranks agree (one all-reduce) whether all of them got through it, so a
    rank that fails never leaves the others waiting in a collective."""
    from fastqandfurious_b200 import shard
    desc, kind, gib, halo = CONFIGS[key]
    gib = {'illumina': args.cfg3_gib, 'ont': args.cfg4_gib, 'multiline': args.cfg5_gib}[kind]
    total = int(gib * (1 << 30))
    out = {'config': desc, 'kind': kind}
    state = {'job': None, 'err': None}

    def phase(fn):
        ok = 1
        try:
            if state['err'] is None:
                fn()
        except Exception as exc:
            import traceback
            traceback.print_exc(file=sys.stderr)
            state['err'] = repr(exc)
        ok = 0 if state['err'] is not None else 1
        if world > 1:
            t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0 and state['err'] is None:
                state['err'] = 'another rank failed'
        return state['err'] is None

    def build():
        state['job'] = shard.SynthJob(kind, total, rank, world, dev, halo_bytes=halo, cfg=args.cfg)
        state['job'].prepare()

    def warm():
        for _ in range(3):
            state['job'].step()
        torch.cuda.synchronize()
        state['job'].read()

    def verify():
        state['verified'] = state['job'].verify_local()
        state['wins'], state['wrecs'] = check_seam_windows(state['job'])

    def timed():
        job = state['job']
        est = max(job.global_bytes() / world / 3e12, 2e-4)  # steps: a stable mean within about a second of device time
        steps = int(max(3, min(args.steps, 1.0 / est)))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            job.step()
        e1.record()
        torch.cuda.synchronize()
        state['ms'], state['steps'] = e0.elapsed_time(e1) / steps, steps
        job.read()  # raises on a device-side error in the timed steps

    try:
        if phase(build) and phase(warm) and phase(verify) and phase(timed):
            job = state['job']
            ms, verified, wins, wrecs = state['ms'], state['verified'], state['wins'], state['wrecs']
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
                t = torch.tensor([verified, wins, wrecs], dtype=torch.int64, device=dev)
                dist.all_reduce(t)
                verified, wins, wrecs = (int(x) for x in t.tolist())
            nbytes, nrec = job.global_bytes(), job.global_records()
            if verified != nrec:
                raise AssertionError('%d rows verified, the stream has %d records' % (verified, nrec))
            alg = nbytes + 48 * nrec
            out.update({
                'ms_per_step': ms, 'steps': state['steps'], 'gbs': nbytes / ms / 1e6, 'mrec_s': nrec / ms / 1e3,
                'bytes': nbytes, 'records': nrec, 'bytes_per_gpu': job.plan.own_len, 'scaling': 'strong',
                'path': ('general (exact)' if job.exact else 'general (speculative pass)') if job.general else 'fast4',
                'roofline': {'bound': 'hbm', 'algorithmic_bytes_per_step': alg, 'achieved': alg / ms / 1e6 / world,
                             'peak': peak, 'unit': 'GB/s per GPU', 'frac': alg / ms / 1e6 / world / peak,
                             'hbm_read_frac': nbytes / ms / 1e6 / world / peak},
                'rows_verified_vs_generator_truth': verified,
                'seam_windows_vs_reference': {'windows': wins, 'records': wrecs,
                                              'what': 'windows of +-max(1 MiB, halo/2) around every shard seam (N=1: the '
                                                      '7 seams of an 8-way split), parsed by oracle/_ref'},
                'sharding': 'none' if world == 1 else '%d byte-range shards cut at arbitrary bytes, %d-byte halo (%s)' % (
                    world, halo, job.parser.transport)})
            # the reference on this box's cores, first 64 MiB of the stream (N=1 only: a reported baseline)
            if world == 1 and not args.no_cpu:
                k_hi = int(torch.searchsorted(job.stream.off, torch.tensor([64 << 20], device=dev), right=True).item()) - 1
                starts = job.stream.off[:k_hi + 1].cpu().numpy()
                sample = job.buf[:int(starts[-1])].cpu().numpy().tobytes()
                r = cpu_reference_rate(sample, os.cpu_count() or 1, 2, starts=starts,
                                       fbufsize=(4 << 20) if kind == 'ont' else 2 ** 16)
                out['cpu_reference'] = {'gbs': r['gbs'], 'mrec_s': r['mrec_s'], 'cores': r['cores'], 'kind': r['kind'],
                                        'sample': 'first %d MiB of the stream, %d record-aligned slices x 2 passes' % (
                                            len(sample) >> 20, r['cores'])}
        else:
            out['failed'] = state['err']
    except Exception as exc:  # an extras entry never takes the headline down
        import traceback
        traceback.print_exc(file=sys.stderr)
        out['failed'] = repr(exc)
    finally:
        if state['job'] is not None:
            state['job'].free()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import fastqandfurious_b200 as fq
    from fastqandfurious_b200 import _lib, device, shard

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa_cpus = device.bind_host_to_gpu(dev) if world > 1 else None  # pinned staging next to the GPU
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=300))
    desc, nbytes = WORKLOADS[args.workload]
    L = _lib.lib()

    # ---- synthetic input, resident in HBM before the timed region --------------------------------
    if world == 1:
        nrec_total = nbytes // REC_BYTES
        buf = fq.synth_fixed(nrec_total)
        job = None
    else:
        # halo: the caller's bound on the largest record (the reference's own contract for fbufsize,
        # src/fastqandfurious.py:219-223)
        job = shard.ShardedJob.synthetic(nbytes, REC_BYTES, rank, world, dev, cfg=args.cfg, halo_bytes=args.halo)
        buf = job.buf
    cap = buf.numel() // REC_BYTES + 64
    table = torch.empty((cap, 6), dtype=torch.int64, device=dev)
    result = torch.empty(16, dtype=torch.int64, device=dev)
    flags = _lib.FLAG_CFG(args.cfg) | _lib.FLAG_FAST_ONLY

    def step():
        if job is None:
            device.parse_raw(buf, 1, -1, table, None, 0, result, flags)
        else:
            job.step(table, result, flags)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    res = device.read_result(result) if job is None else job.result()
    assert res.error == 0 and not res.need_general, (res.error, res.need_general, res.first_bad)
    nrec_step = res.n_records if job is None else job.records_per_step()
    if job is not None:  # closed-form truth of the synthetic stream: every row of every shard, and no gaps between shards
        verified_records = job.verify_fixed()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # the sampler's start-up (a subprocess and a 0.3 s pause) leaves the GPU idle: a few untimed steps bring the clocks
    # back up before the timed region (without them the first milliseconds of the 47 ms region run slow: 0.2435 against
    # 0.2340 ms per step in otherwise identical runs)
    for _ in range(max(args.warmup, 3) * 4):
        step()
    launches0 = device.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = device.launch_count - launches0
    # the dominant kernel's duration: the same K steps again, this time with a CUDA event on either side of every
    # fq_scan_kernel launch (fqb_profile_*).  Kept out of the region above because an event record between two kernels
    # of a step costs about 2.5 us each (0.2414 against 0.2361 ms per step, profiles/r02_*): `value` is the step as a
    # caller runs it, `ms_per_step_with_kernel_events` the same loop with the events in it.
    import ctypes
    _lib.check(L.fqb_profile_enable(1), 'fqb_profile_enable')
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record()
    for _ in range(args.steps):
        step()
    pv1.record()
    torch.cuda.synchronize()
    ms_prof = pv0.elapsed_time(pv1)
    tot = ctypes.c_double()
    cnt = ctypes.c_int64()
    _lib.check(L.fqb_profile_read(ctypes.byref(tot), ctypes.byref(cnt)), 'fqb_profile_read')
    L.fqb_profile_enable(0)
    if world > 1:
        t = torch.tensor([ms, ms_prof], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_prof = float(t[0].item()), float(t[1].item())
    sampler.window(t0, t1)
    clocks = sampler.stop() if rank == 0 else None

    bytes_step = buf.numel() * world if job is None else job.global_bytes()
    recs_step = nrec_step if job is None else job.global_records()
    value = bytes_step * args.steps / (ms / 1e3) / 1e9

    # ---- end to end through the host-buffer API --------------------------------------------------
    e2e = None
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    sharded_e2e = (job is not None and job.parser.transport == 'fused' and not args.e2e_independent)
    e2e_extra = {}
    if sharded_e2e:
        # N > 1: ONE logical stream; rank g's byte range sits in its pinned host memory, the shards are stitched on the
        # devices (halo pulled from the right neighbour, line counts from the left ones) and every rank gets the rows of
        # the records it owns back in host memory (shard.ShardedHostParser, double-buffered: the copies of step k + 1
        # overlap the parse of step k)
        plan = job.parser.plan
        shp = shard.ShardedHostParser(plan, dev, cfg=args.cfg, cap=plan.own_len // REC_BYTES + 64)
        ebuf = job.parser.own(0)
        host = torch.empty(plan.own_len, dtype=torch.uint8).pin_memory()
        host.copy_(ebuf)
        torch.cuda.synchronize()
        for _ in range(2):
            k0_rows = shp.parse(host)
        dist.barrier()
        torch.cuda.synchronize()
        te0 = time.perf_counter()
        shp.submit(host)
        for _ in range(e2e_steps - 1):
            shp.submit(host)
            k0_rows = shp.collect()
        k0_rows = shp.collect()
        torch.cuda.synchronize()
        te = time.perf_counter() - te0
        t = torch.tensor([te], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
        # the last step's rows against the closed form of the synthetic stream, and the ranks' record ranges must tile it
        import numpy as np
        k0, rows, _res = k0_rows
        k = np.arange(k0, k0 + len(rows), dtype=np.int64) * REC_BYTES
        want = np.stack([k, k + 32, k + 33, k + 183, k + 186, k + 336], axis=1)
        rows_ok = bool(np.array_equal(rows, want))
        every = torch.empty(2 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, torch.tensor([k0, len(rows) if rows_ok else -1], dtype=torch.int64, device=dev))
        nxt, tiled = 0, True
        for a_, c_ in every.view(-1, 2).tolist():
            tiled = tiled and a_ == nxt and c_ >= 0
            nxt = a_ + c_
        e2e_records = nxt if tiled else -1
        h2d_b, d2h_b = plan.total, 0
        t = torch.tensor([shp.stats['d2h_bytes'] / max(1, shp.stats['h2d_bytes'] // plan.own_len)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        d2h_b = int(t.item())
        e2e_api = ('fastqandfurious_b200.shard.ShardedHostParser.submit/collect (one stream, rank g holds bytes [c_g, c_g + %d) '
                   'in pinned host memory -> (first record index, int64[n,6] rows) per rank in host memory)' % plan.own_len)
        e2e_extra = {'sharded': True, 'rows_verified': e2e_records, 'halo_bytes': plan.halo_bytes,
                     'stitching': 'halo pulled from the right neighbour over NVLink, line counts published by the left ones'}
        e2e_bytes = plan.total
        del shp
    else:
        hp = device.HostParser(dev, chunk_bytes=args.e2e_chunk, cfg=args.cfg)
        # N > 1 without peer memory: every rank streams its own record-aligned host buffer (independent streams)
        ebuf = buf if job is None else fq.synth_fixed(buf.numel() // REC_BYTES, device=dev)
        host = torch.empty(ebuf.numel(), dtype=torch.uint8).pin_memory()
        host.copy_(ebuf)
        torch.cuda.synchronize()
        for _ in range(2):
            rows = hp.parse(host)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        te0 = time.perf_counter()
        for _ in range(e2e_steps):
            rows = hp.parse(host)
        torch.cuda.synchronize()
        te = time.perf_counter() - te0
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        h2d_b, d2h_b = hp.stats['h2d_bytes'] * world, hp.stats['d2h_bytes'] * world
        e2e_records = int(len(rows)) * world
        e2e_api = 'fastqandfurious_b200.device.HostParser.parse (pinned host tensor -> int64[n,6] host table)'
        e2e_extra = {'sharded': False, 'chunk_bytes': args.e2e_chunk}
        e2e_bytes = host.numel() * world
        del hp
    # the link itself: the same pinned buffer copied host -> device with nothing else going on (this rank)
    link_ms = None
    try:
        sink = torch.empty_like(ebuf)
        for _ in range(2):
            sink.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            sink.copy_(host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link_ms = c0.elapsed_time(c1) / 3
        del sink
    except Exception:
        link_ms = None
    e2e = {'value': e2e_bytes * e2e_steps / te / 1e9, 'unit': 'GB/s',
           'h2d_bytes_per_step': h2d_b, 'd2h_bytes_per_step': d2h_b,
           'steps': e2e_steps, 'records': e2e_records, 'api': e2e_api,
           'host_cpus_bound': len(numa_cpus) if numa_cpus else None,
           'h2d_link_gbs_one_gpu': (host.numel() / link_ms / 1e6) if link_ms else None}
    e2e.update(e2e_extra)

    # ---- N > 1: the sharded parse with Phred decode (fqb_shard_scan_decode), not the headline ------------------
    shard_extras = None
    if job is not None and not args.no_extras:
        dec_ms, dec_ok = -1.0, 0
        qual = None
        try:
            qual = job.parser.alloc_qual()
        except Exception as exc:
            print('bench: no memory for the Phred mirror on rank %d: %r' % (rank, exc), file=sys.stderr)
        ready = torch.tensor([1.0 if qual is not None else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ready, op=dist.ReduceOp.MIN)  # every rank steps or none does: a shard's emit waits for its peers
        try:  # no collective call in here: a rank that fails must not leave the others waiting
            if float(ready.item()) < 1.0:
                raise RuntimeError('a rank could not allocate the Phred mirror')
            dsteps = max(3, min(args.steps, 50))
            for _ in range(3):
                job.parser.step(job.table, qual=qual)
            torch.cuda.synchronize()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for _ in range(dsteps):
                job.parser.step(job.table, qual=qual)
            d1.record()
            torch.cuda.synchronize()
            dec_ms = d0.elapsed_time(d1) / dsteps
            r = job.parser.read()
            rows = job.table[:min(int(r.n_records), 1 << 16)]
            idx = (rows[:, 4] - job.parser.plan.offset).unsqueeze(1) + torch.arange(150, device=dev)
            want_q = (job.parser.buf[idx.reshape(-1)].to(torch.int16) - 33).to(torch.int8)
            dec_ok = 1 if (len(rows) > 0 and torch.equal(qual[idx.reshape(-1)], want_q)
                           and bool((rows[:, 5] - rows[:, 4] == 150).all())) else 0
            del qual
        except Exception as exc:
            print('bench: sharded decode extra failed on rank %d: %r' % (rank, exc), file=sys.stderr)
        t = torch.tensor([dec_ms, -float(dec_ok)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if float(t[1].item()) == -1.0 and float(t[0].item()) > 0:
            dms = float(t[0].item())
            shard_extras = {'sharded_with_phred_decode': {
                'ms_per_step': dms, 'gbs': job.global_bytes() / (dms / 1e3) / 1e9,
                'quality_strings_checked_per_gpu': int(min(int(nrec_step), 1 << 16)),
                'api': 'ShardedParser.step(table, qual=alloc_qual()) -> fqb_shard_scan_publish_ready(d_qual) + fqb_shard_emit_wait'}}

    # ---- the other BASELINE configs (3: 64 GiB Illumina-like, 4: ONT-like long reads, 5: wrapped multi-line), every
    #      rank takes part; the headline's buffers go first (64 GiB + lists + table need the room) -----------------
    peak, peak_src = measured_peak_gbs()
    host_bytes_1g = None
    if world == 1 and rank == 0 and not (args.no_cpu and args.no_extras):
        host_bytes_1g = host.numpy().tobytes()  # the workload's bytes for the drop-in / CPU legs below
    headline = {'buf_numel': buf.numel(), 'transport': job.parser.transport if job is not None else None,
                'info': device.kernel_info(args.cfg)}
    extras_1gpu = None
    if world == 1 and not args.no_extras:
        extras_1gpu = measure_extras(fq, device, _lib, torch, buf, table, args)
    del host, ebuf, table, buf
    job = None
    device._ws_cache.clear()
    torch.cuda.empty_cache()
    cfg_extras = {}
    if not args.no_configs:
        for key in CONFIGS:
            if key.startswith('cfg5') and world > 1 and not args.cfg5_sharded:
                continue  # BASELINE.json runs config 5 on one GPU
            cfg_extras[key] = measure_config(key, args, rank, world, dev, torch, dist, peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (this rank) ---------------------------------------------
    # fq_scan_kernel is the only kernel that touches the input: its algorithmic bytes per launch are
    # the input bytes (337 B per 150 bp record x records per launch).  The 48 B/record table
    # is written by fq_emit_kernel; `pipeline` below is the whole step against the whole algorithmic
    # volume (337 + 48 B per record).
    alg_bytes = headline['buf_numel']
    scan_ms = tot.value / max(1, cnt.value)
    achieved = alg_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'scan_traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    info = headline['info']
    step_ms = ms / args.steps
    pipe_bytes = alg_bytes + 48 * nrec_step
    roofline = {'bound': 'hbm', 'kernel': 'fq_scan_kernel', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak if achieved else None, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': scan_ms, 'kernel_launches_timed': cnt.value,
                'kernel_share_of_step': scan_ms / (ms_prof / args.steps) if ms_prof > 0 else None,
                'ms_per_step_with_kernel_events': ms_prof / args.steps, 'kernel_config': info,
                'pipeline': {'algorithmic_bytes_per_step': pipe_bytes, 'achieved': pipe_bytes / (step_ms / 1e3) / 1e9,
                             'frac': pipe_bytes / (step_ms / 1e3) / 1e9 / peak,
                             'kernels_per_step': ['memset(state)', 'fq_scan_kernel', 'fq_emit_kernel']}}

    # ---- other rows of the scope table on one GPU (not the headline; same timing method) ----------
    extras = dict(shard_extras or {})
    extras.update(extras_1gpu or {})
    extras.update(cfg_extras)

    # ---- CPU baseline: the reference's C extension on this box's cores (bounded sample) ----------
    cpu = None
    if world == 1 and not args.no_cpu:
        r, sample_desc, sample = cpu_baseline_block(min(nbytes, args.ref_sample), args.cpu_bytes)
        cpu = {'value': r['gbs'], 'unit': 'GB/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': sample_desc,
               'mrecords_per_s': r['mrec_s'], 'cpu_model': cpu_model(), 'host_cpus': os.cpu_count(),
               'single_process': cpu_single_process_legs(sample)}
        # the reference's own call, per record, through OUR drop-in: next to single_process.readfastq_iter_abspos_*
        try:
            e2e['dropin_per_record'] = dropin_rate(fq, host_bytes_1g if host_bytes_1g is not None else sample)
        except Exception as exc:
            e2e['dropin_per_record'] = {'failed': repr(exc)}
        del sample

    line = {
        'metric': METRIC, 'value': value, 'unit': 'GB/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
        'data': 'synthetic', 'mrecords_per_s': recs_step * args.steps / (ms / 1e3) / 1e6,
        'config': static_config(args.workload),
        'run': {'sharded_rows_verified': None if world == 1 else int(verified_records),
                'sharding': 'none' if world == 1 else 'byte-range shards of one stream, %d-byte halo, neighbour exchange (%s)' % (args.halo, headline['transport']),
                'timed_step': 'memset(state) + fq_scan_kernel + fq_emit_kernel (programmatic dependent launch) enqueued back to back (FQB_FLAG_FAST_ONLY); the '
                              'result header stays on the device -- the host read-back every product call pays '
                              '(device.read_result, one 128-byte D2H + sync) is inside e2e, not inside value'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu,
        'extras': extras,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='fixed150_1g', choices=sorted(WORKLOADS))
    ap.add_argument('--cfg', type=int, default=0, help='scan kernel configuration (tuning)')
    ap.add_argument('--halo', type=int, default=1 << 20, help='halo bytes per shard at N>1 (>= the largest record)')
    ap.add_argument('--e2e-steps', type=int, default=10)
    ap.add_argument('--e2e-chunk', type=int, default=1 << 26)
    ap.add_argument('--e2e-independent', action='store_true', help='N > 1: independent per-rank streams instead of the sharded one')
    ap.add_argument('--cpu-bytes', type=float, default=float(4 << 30), help='bytes the CPU baseline parses in total')
    ap.add_argument('--ref-sample', type=int, default=1 << 30, help='bytes of the workload the CPU reference parses per pass')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip BASELINE configs 3-5 (extras.cfg*)')
    ap.add_argument('--cfg3-gib', type=float, default=64.0, help='GiB of the whole Illumina-like stream (strong scaling)')
    ap.add_argument('--cfg4-gib', type=float, default=8.0, help='GiB of the whole ONT-like stream')
    ap.add_argument('--cfg5-gib', type=float, default=8.0, help='GiB of the whole multi-line stream')
    ap.add_argument('--cfg5-sharded', action='store_true', help='also run config 5 (general path) sharded at N > 1')
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print meanwhile (NCCL banner, make) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            if args.impl == 'reference':
                run_reference(args)
            else:
                run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    line = out.getvalue().strip()
    if line:
        print(line.splitlines()[-1], flush=True)


if __name__ == '__main__':
    main()
