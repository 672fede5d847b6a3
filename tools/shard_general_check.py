"""torchrun check of the sharded GENERAL path on real GPUs: multi-line input cut into N byte-range shards, fast
path first (it must decline), then step_general on every rank; rank 0 compares the stitched rows with the
single-buffer parse of the whole stream and prints the time of the sharded general step."""
import os
import sys
sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import numpy as np
import torch
import torch.distributed as dist
import fqgen
import fastqandfurious_b200 as fq
from fastqandfurious_b200 import device, shard

world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
base = fqgen.variable_records_np(120000, 31, 'multiline')
reps = int(os.environ.get('REPS', '4')) * world
data = np.tile(base, reps)
total = len(data)
cut = [total * g // world + (7 * g) % 13 for g in range(world)] + [total]
own_lens = [cut[g + 1] - cut[g] for g in range(world)]
plan = shard.ShardPlan(rank, world, own_lens, 1 << 20)
P = shard.ShardedParser(plan, dev)
P.own().copy_(torch.from_numpy(data[cut[rank]:cut[rank + 1]].copy()).to(dev))
table = torch.empty((plan.own_len // 64 + 4096, 6), dtype=torch.int64, device=dev)
P.step(table)
assert P.needs_general(), 'the fast path should have declined this input'
for _ in range(2):
    P.step_general(table)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
P.step_general(table)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
res = device.read_result(P.result)
assert res.error == 0, res.error
n = torch.tensor([res.n_records], dtype=torch.int64, device=dev)
counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
dist.all_gather(counts, n)
counts = [int(c.item()) for c in counts]
assert res.reserved[0] == sum(counts[:rank]), (rank, res.reserved[0], counts)
if rank == 0:
    parts = [table[:counts[0]].clone()]
    for g in range(1, world):
        t = torch.empty((counts[g], 6), dtype=torch.int64, device=dev)
        dist.recv(t, src=g)
        parts.append(t)
    got = torch.cat(parts)
    whole = fq.parse_buffer(torch.from_numpy(data).to(dev), cap=len(data) // 64 + 4096)
    assert whole.path == 2
    assert torch.equal(got, whole.table), (got.shape, whole.table.shape)
    print('sharded general path on %d GPUs: %d records identical to the single-buffer parse; %.3f ms for %.2f GiB = %.1f GB/s'
          % (world, got.shape[0], ms.item(), total / 2 ** 30, total / ms.item() / 1e6), flush=True)
else:
    dist.send(table[:counts[rank]].contiguous(), dst=0)
dist.barrier()
dist.destroy_process_group()
