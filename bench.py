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
def _cpu_worker(args):
    data, repeats, use_ref = args
    import io
    from array import array
    import oracle
    nrec = 0
    t0 = time.perf_counter()
    if use_ref:
        mod, cext = oracle.reference()
        for _ in range(repeats):
            out = array('q')
            for pos in mod.readfastq_iter(io.BytesIO(data), 2 ** 16, entryfunc=mod.entryfunc_abspos,
                                          entrypos=cext.entrypos):
                out.extend(pos)
            nrec += len(out) // 6
    else:
        for _ in range(repeats):
            table, err, _ = oracle.readfastq(data)
            nrec += len(table)
    return time.perf_counter() - t0, nrec, len(data) * repeats


def cpu_reference_rate(sample, procs, repeats):
    """Whole-file GB/s of the reference's own path (readfastq_iter + C entrypos + entryfunc_abspos,
    fbufsize 2**16 as in src/demo/benchmark.py:26-27) over `sample` (bytes), split into `procs`
    record-aligned slices, one process each (the reference itself is single-threaded)."""
    import multiprocessing as mp
    import oracle
    oracle.build()
    use_ref = oracle.reference() is not None
    nrec_total = len(sample) // REC_BYTES
    per = max(1, nrec_total // procs)
    slices = [bytes(sample[i * per * REC_BYTES:(i + 1) * per * REC_BYTES]) for i in range(procs)]
    slices = [s for s in slices if s]
    ctx = mp.get_context('fork')
    t0 = time.perf_counter()
    if len(slices) == 1:
        results = [_cpu_worker((slices[0], repeats, use_ref))]
    else:
        with ctx.Pool(len(slices)) as pool:
            results = pool.map(_cpu_worker, [(s, repeats, use_ref) for s in slices])
    wall = time.perf_counter() - t0
    nbytes = sum(r[2] for r in results)
    nrec = sum(r[1] for r in results)
    worker_wall = max(r[0] for r in results)
    return {'gbs': nbytes / worker_wall / 1e9, 'mrec_s': nrec / worker_wall / 1e6, 'bytes': nbytes, 'wall_s': wall,
            'cores': len(slices), 'kind': 'reference' if use_ref else 'port'}


def host_sample(nbytes):
    """First `nbytes` of the workload on the host (numpy twin of the device generator)."""
    import fqgen
    return fqgen.fixed_records_np(nbytes // REC_BYTES).tobytes()


def cpu_baseline_block(sample_bytes, total_bytes):
    procs = os.cpu_count() or 1
    sample = host_sample(sample_bytes)
    repeats = max(1, int(total_bytes // max(1, len(sample))))
    r = cpu_reference_rate(sample, procs, repeats)
    return r, ('first %d MiB of the workload, %d record-aligned slices x %d passes = %.1f GiB parsed, '
               'readfastq_iter(fbufsize=65536)+C entrypos+entryfunc_abspos' %
               (sample_bytes >> 20, r['cores'], repeats, r['bytes'] / 2 ** 30))


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    desc, nbytes = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    sample = host_sample(min(nbytes, 256 << 20))
    # each step parses the sample once per core-slice; bounded so K+W steps end within minutes
    times, nrec = [], 0
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
        'config': {'workload': args.workload, 'description': desc, 'record_bytes': REC_BYTES},
        'cpu_baseline': {'value': gbs, 'unit': 'GB/s', 'cores': r['cores'], 'kind': r['kind'],
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
            gflags = _lib.FLAG_CFG(args.cfg) | _lib.FLAG_FORCE_GENERAL
            ml = res.n_lines + 64
            ms = _time_steps(torch, lambda: device.parse_raw(d, 1, -1, tab, None, 0, result, gflags, max_lines=ml), steps)
        out[name] = {'gbs': d.numel() / ms / 1e6, 'ms_per_step': ms, 'path': 'fast4' if res.path == 1 else 'general',
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
        dist.init_process_group('nccl', device_id=dev)
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
    _lib.check(L.fqb_profile_enable(1), 'fqb_profile_enable')
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
    import ctypes
    tot = ctypes.c_double()
    cnt = ctypes.c_int64()
    _lib.check(L.fqb_profile_read(ctypes.byref(tot), ctypes.byref(cnt)), 'fqb_profile_read')
    L.fqb_profile_enable(0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sampler.window(t0, t1)
    clocks = sampler.stop() if rank == 0 else None

    bytes_step = buf.numel() * world if job is None else job.global_bytes()
    recs_step = nrec_step if job is None else job.global_records()
    value = bytes_step * args.steps / (ms / 1e3) / 1e9

    # ---- end to end through the host-buffer API --------------------------------------------------
    e2e = None
    hp = device.HostParser(dev, chunk_bytes=args.e2e_chunk, cfg=args.cfg)
    # at N>1 every rank streams its own record-aligned host buffer of the same size (independent streams:
    # host memory is per process, there is nothing to stitch)
    ebuf = buf if job is None else fq.synth_fixed(buf.numel() // REC_BYTES, device=dev)
    host = torch.empty(ebuf.numel(), dtype=torch.uint8).pin_memory()
    host.copy_(ebuf)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
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
    e2e = {'value': host.numel() * world * e2e_steps / te / 1e9, 'unit': 'GB/s',
           'h2d_bytes_per_step': hp.stats['h2d_bytes'] * world, 'd2h_bytes_per_step': hp.stats['d2h_bytes'] * world,
           'steps': e2e_steps, 'records': int(len(rows)) * world, 'chunk_bytes': args.e2e_chunk,
           'api': 'fastqandfurious_b200.device.HostParser.parse (pinned host tensor -> int64[n,6] host table)',
           'host_cpus_bound': len(numa_cpus) if numa_cpus else None,
           'h2d_link_gbs_one_gpu': (host.numel() / link_ms / 1e6) if link_ms else None}

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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (this rank) ---------------------------------------------
    # fq_scan_kernel is the only kernel that touches the input: its algorithmic bytes per launch are
    # the input bytes (337 B per 150 bp record x records per launch).  The 48 B/record table
    # is written by fq_emit_kernel; `pipeline` below is the whole step against the whole algorithmic
    # volume (337 + 48 B per record).
    peak, peak_src = measured_peak_gbs()
    alg_bytes = buf.numel()
    scan_ms = tot.value / max(1, cnt.value)
    achieved = alg_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'scan_traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    info = device.kernel_info(args.cfg)
    step_ms = ms / args.steps
    pipe_bytes = buf.numel() + 48 * nrec_step
    roofline = {'bound': 'hbm', 'kernel': 'fq_scan_kernel', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak if achieved else None, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': scan_ms, 'kernel_launches_timed': cnt.value,
                'kernel_share_of_step': scan_ms / step_ms if ms > 0 else None, 'kernel_config': info,
                'pipeline': {'algorithmic_bytes_per_step': pipe_bytes, 'achieved': pipe_bytes / (step_ms / 1e3) / 1e9,
                             'frac': pipe_bytes / (step_ms / 1e3) / 1e9 / peak,
                             'kernels_per_step': ['memset(state)', 'fq_scan_kernel', 'fq_emit_kernel']}}

    # ---- other rows of the scope table on one GPU (not the headline; same timing method) ----------
    extras = shard_extras
    if world == 1 and not args.no_extras:
        extras = measure_extras(fq, device, _lib, torch, buf, table, args)

    # ---- CPU baseline: the reference's C extension on this box's cores (bounded sample) ----------
    cpu = None
    if world == 1 and not args.no_cpu:
        r, sample_desc = cpu_baseline_block(min(nbytes, 256 << 20), args.cpu_bytes)
        cpu = {'value': r['gbs'], 'unit': 'GB/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': sample_desc,
               'mrecords_per_s': r['mrec_s']}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'GB/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
        'data': 'synthetic', 'mrecords_per_s': recs_step * args.steps / (ms / 1e3) / 1e6,
        'config': {'workload': args.workload, 'description': desc, 'bytes_per_gpu': buf.numel(),
                   'records_per_gpu': int(nrec_step), 'record_bytes': REC_BYTES,
                   'l2': 'input (1 GiB/GPU) is larger than L2 (126 MB); no flush needed',
                   'sharded_rows_verified': None if world == 1 else int(verified_records),
                   'sharding': 'none' if world == 1 else 'byte-range shards of one stream, %d-byte halo, neighbour exchange (%s)' % (args.halo, job.parser.transport)},
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
    ap.add_argument('--cpu-bytes', type=float, default=float(4 << 30), help='bytes the CPU baseline parses in total')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
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
