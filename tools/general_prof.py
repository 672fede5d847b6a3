import sys, os
sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import torch, fqgen
import fastqandfurious_b200 as fq
from fastqandfurious_b200 import device, _lib
base = fqgen.variable_records_np(120000, 31, 'multiline')
reps = (1 << 30) // len(base)
d = torch.from_numpy(base.copy()).cuda().repeat(reps)
res = fq.parse_buffer(d, cap=120000 * reps + 64)
print('path', res.path, 'n', res.n, 'lines', res.n_lines)
tab = res.table_full
result = torch.empty(16, dtype=torch.int64, device='cuda')
for _ in range(3):
    device.parse_raw(d, 1, -1, tab, None, 0, result, _lib.FLAG_FORCE_GENERAL, max_lines=res.n_lines + 64)
torch.cuda.synchronize()
