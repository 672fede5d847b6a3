"""Per-phase device times of one sharded step (torchrun, N ranks), double-buffered shards (ShardedParser(double_buffer=
True), the bench's headline): on the main stream  wait for the halo | scan + count/publish/signal | emit (waits for the
counts)  and, on the pull stream, ready-wait + peer copy of the next parse's halo.  Prints one line per rank."""
import ctypes
import os
import statistics
import sys

sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import torch
import torch.distributed as dist

from fastqandfurious_b200 import _lib, device, shard

world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
job = shard.ShardedJob.synthetic(1 << 30, 337, rank, world, dev, double_buffer=True)
P = job.parser
plan, L = P.plan, _lib.lib()
assert P.double, 'needs the fused transport'
for _ in range(10):
    job.step()
torch.cuda.synchronize()
dist.barrier()
STEPS = 60
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(STEPS)]
pv = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(STEPS)]
n, own = P.n, plan.own_len
cur = torch.cuda.current_stream()
for it in range(STEPS):
    e = ev[it]
    k = P.epoch + 1
    b = k % 2
    buf = P.bufs[b]
    e[0].record()
    cur.wait_event(P.pull_done[b])
    e[1].record()
    stream = ctypes.c_void_p(cur.cuda_stream)
    sentinel = 1 if rank == 0 else 0
    P.epoch = k
    par = k % shard.SLOT_RING
    sig = P.ready_left is not None
    _lib.check(L.fqb_shard_scan_publish_ready(buf.data_ptr(), n, own, sentinel, P.own_lines.data_ptr(), P.pub_ptrs[par], P.n_pub,
                                              k, P.ready_left if sig else None, P._signalled + 1 if sig else 0, None, 0,
                                              P.ws.data_ptr(), P.ws.numel(), P.flags, stream), 'scan')
    if sig:
        P._signalled += 1
    e[2].record()
    wait = P.slots.data_ptr() + par * plan.world * 2 * 8
    _lib.check(L.fqb_shard_emit_wait(buf.data_ptr(), n, own, sentinel, 1 if plan.is_last else 0, plan.offset - sentinel, wait,
                                     plan.rank, k, job.table.data_ptr(), job.table.shape[0], P.result.data_ptr(),
                                     P.ws.data_ptr(), P.ws.numel(), P.flags, stream), 'emit')
    e[3].record()
    done = torch.cuda.Event()
    done.record(cur)
    P.parse_done[b] = done
    # the next parse's halo, on the pull stream, bracketed by events of that stream
    kb = (k + 1) % 2
    P.pull_stream.wait_event(P.parse_done[kb]) if P.parse_done[kb] is not None else None
    with torch.cuda.stream(P.pull_stream):
        pv[it][0].record(P.pull_stream)
        if plan.halo_len():
            _lib.check(L.fqb_shard_wait_ready(P.ready.data_ptr(), k + 1, P.halo_status.data_ptr(),
                                              ctypes.c_void_p(P.pull_stream.cuda_stream)), 'wait')
            P.bufs[kb][own:own + plan.halo_len()].copy_(P.rights[kb][:plan.halo_len()], non_blocking=True)
        pv[it][1].record(P.pull_stream)
        P.pull_done[kb].record(P.pull_stream)
torch.cuda.synchronize()
res = P.read()
seg = [[ev[it][q].elapsed_time(ev[it][q + 1]) * 1e3 for it in range(10, STEPS)] for q in range(3)]
tot = [ev[it][0].elapsed_time(ev[it + 1][0]) * 1e3 for it in range(10, STEPS - 1)]
pull = [pv[it][0].elapsed_time(pv[it][1]) * 1e3 for it in range(10, STEPS)]
med = statistics.median
print('rank %d of %d: wait for halo %.1f us | scan + count/publish/signal %.1f us | emit (incl. wait for counts) %.1f us | '
      'step %.1f us (with the events of this tool) || pull stream: ready-wait + %d-byte peer copy %.1f us; records %d' %
      (rank, world, med(seg[0]), med(seg[1]), med(seg[2]), med(tot), plan.halo_len(), med(pull), res.n_records), flush=True)
dist.barrier()
dist.destroy_process_group()
