"""torchrun check of the sharded FASTA parse on real GPUs (NCCL): one seeded stream, every rank takes its byte range
to its device, shard.ShardedFastaParser exchanges look-behind / halo bytes and counts, every rank compares its rows
with its slice of the single-buffer parse of the whole stream (done on the same device) and rank 0 prints the time."""
import os
import sys
sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import numpy as np
import torch
import torch.distributed as dist
from fastqandfurious_b200 import device, shard

world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rng = np.random.default_rng(6)
nrec = 60000
rec = bytearray()
lens = rng.integers(0, 900, nrec)
for k in range(nrec):
    rec += b'>read%07d sample %d\n' % (k, lens[k])
    s = rng.choice(np.frombuffer(b'ACGT', dtype=np.uint8), size=int(lens[k])).tobytes()
    rec += b'\n'.join(s[i:i + 60] for i in range(0, len(s), 60)) + b'\n'
data = np.tile(np.frombuffer(bytes(rec), dtype=np.uint8), int(os.environ.get('REPS', '8')))
total = len(data)
cut = [total * g // world + (7 * g) % 13 for g in range(world)] + [total]
own_lens = [cut[g + 1] - cut[g] for g in range(world)]
sp = shard.ShardedFastaParser(rank, world, own_lens, halo_bytes=1 << 20, lookbehind_bytes=1 << 16)
own = torch.from_numpy(data[cut[rank]:cut[rank + 1]].copy()).to(dev)
k0, rows, n, status, tail_pos, resume = sp.parse(own)  # warm-up and the answer
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    sp.parse(own)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
whole = device.parse_fasta_buffer(torch.from_numpy(data).to(dev))
ok = (n == whole.n and status == whole.tail_status and tail_pos == whole.tail_pos and resume == whole.resume_offset and
      torch.equal(rows, whole.table[k0:k0 + rows.shape[0]]))
flag = torch.tensor([1 if ok else 0, rows.shape[0]], dtype=torch.int64, device=dev)
dist.all_reduce(flag)
if rank == 0:
    print('sharded FASTA over %d GPUs: %d bytes, %d records, all ranks identical to the single-buffer parse: %s '
          '(rows summed over ranks %d), %.3f ms per parse incl. the exchange and the host round trips = %.1f GB/s'
          % (world, total, n, bool(flag[0].item() == world), int(flag[1].item()), ms.item(), total / ms.item() / 1e6), flush=True)
dist.destroy_process_group()
sys.exit(0 if flag[0].item() == world else 1)
