"""Per-phase device times of one sharded step (torchrun, N ranks): barrier + halo copy | scan + publish | emit (wait)."""
import os
import sys
import ctypes
sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import torch
import torch.distributed as dist
from fastqandfurious_b200 import _lib, device, shard

world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
job = shard.ShardedJob.synthetic(1 << 30, 337, rank, world, dev)
P = job.parser
plan, L = P.plan, _lib.lib()
for _ in range(5):
    job.step()
torch.cuda.synchronize()
dist.barrier()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(50)]
n, own = P.n, plan.own_len
for it in range(50):
    e = ev[it]
    e[0].record()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if P.epoch == 0:
        P.signal_ready()
    _lib.check(L.fqb_shard_pull_halo(P.buf.data_ptr() + own, P.right_ptr, plan.halo_len(), P.ready.data_ptr(), None,
                                     P.epoch + 1, P.halo_status.data_ptr(), stream), 'pull')
    e[1].record()
    sentinel = 1 if rank == 0 else 0
    P.epoch += 1
    par = P.epoch % shard.SLOT_RING
    _lib.check(L.fqb_shard_scan_publish(P.buf.data_ptr(), n, own, sentinel, P.own_lines.data_ptr(), P.pub_ptrs[par], P.n_pub,
                                        P.epoch, P.ws.data_ptr(), P.ws.numel(), P.flags, stream), 'scan')
    P.signal_ready()
    e[2].record()
    wait = P.slots.data_ptr() + par * plan.world * 2 * 8
    _lib.check(L.fqb_shard_emit_wait(P.buf.data_ptr(), n, own, sentinel, 1 if plan.is_last else 0, plan.offset - sentinel, wait,
                                     plan.rank, P.epoch, job.table.data_ptr(), job.table.shape[0], P.result.data_ptr(),
                                     P.ws.data_ptr(), P.ws.numel(), P.flags, stream), 'emit')
    e[3].record()
torch.cuda.synchronize()
# the pull kernel alone (the neighbour's flag is already there)
pe = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
pe[0].record()
for k in range(20):
    _lib.check(L.fqb_shard_pull_halo(P.buf.data_ptr() + own, P.right_ptr, plan.halo_len(), P.ready.data_ptr(), None,
                                     P.epoch, P.halo_status.data_ptr(), stream), 'pull')
    pe[k + 1].record()
torch.cuda.synchronize()
print('rank %d: pull alone %.1f us (halo %d bytes)' % (rank, min(pe[k].elapsed_time(pe[k + 1]) for k in range(20)) * 1e3,
                                                       plan.halo_len()), flush=True)
import statistics
seg = [[ev[it][k].elapsed_time(ev[it][k + 1]) * 1e3 for it in range(5, 50)] for k in range(3)]
tot = [ev[it][0].elapsed_time(ev[it + 1][0]) * 1e3 for it in range(5, 49)]
print('rank %d: barrier+halo %.1f us | scan+publish %.1f us | emit(wait) %.1f us | step %.1f us' %
      (rank, statistics.median(seg[0]), statistics.median(seg[1]), statistics.median(seg[2]), statistics.median(tot)), flush=True)
dist.barrier()
dist.destroy_process_group()
