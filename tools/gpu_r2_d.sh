#!/bin/bash
# round 2, call D: the speculative general pass -- parity tests, fuzz, timing on 1 GiB multi-line, ncu
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log
FUZZ_SECONDS=${FUZZ_SECONDS:-90} timeout -s KILL 600 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -4 gpurun_out/fuzz.log
python - <<'PY' 2>&1 | tail -20
import sys, os
sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import torch, fqgen
import fastqandfurious_b200 as fq
from fastqandfurious_b200 import device, _lib, shard
def t(fn, steps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
job = shard.SynthJob('multiline', 1 << 30, 0, 1, 'cuda')
job.step(); torch.cuda.synchronize()
print('verify', job.verify_local(), 'spec', device.read_result(job.result).reserved[1])
ms = t(job.step)
print('multiline 1 GiB general (spec): %.4f ms  %.1f GB/s' % (ms, job.global_bytes() / ms / 1e6))
flags = _lib.FLAG_FORCE_GENERAL | _lib.FLAG_NO_SPEC
ms = t(lambda: device.parse_raw(job.buf, 1, -1, job.table, None, 0, job.result, flags, max_lines=job.max_lines))
print('multiline 1 GiB general (exact): %.4f ms  %.1f GB/s' % (ms, job.global_bytes() / ms / 1e6))
ms = t(lambda: device.parse_raw(job.buf, 1, -1, job.table, None, 0, job.result, _lib.FLAG_FAST_ONLY))
print('scan+emit only: %.4f ms' % ms)
job.free()
job = shard.SynthJob('illumina', 1 << 30, 0, 1, 'cuda', general=True)
job.step(); torch.cuda.synchronize()
print('verify', job.verify_local(), 'spec', device.read_result(job.result).reserved[1])
ms = t(job.step)
print('illumina 1 GiB forced general (spec): %.4f ms  %.1f GB/s' % (ms, job.global_bytes() / ms / 1e6))
PY
cat > /tmp/gspec_prof.py <<'PY'
import sys
sys.path[:0] = ['.', 'fastq-and-furious_b200', 'tests']
import torch
from fastqandfurious_b200 import shard
job = shard.SynthJob('multiline', 1 << 30, 0, 1, 'cuda')
for _ in range(3): job.step()
torch.cuda.synchronize()
PY
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fq_gspec -s 2 -c 1 -o gpurun_out/prof_gspec python /tmp/gspec_prof.py > gpurun_out/ncu_gspec.log 2>&1; echo "ncu exit $?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gspec.csv python /tmp/gspec_prof.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_gspec.csv')) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r)
out=[]
for r in rows:
    if r is hdr or len(r)!=len(hdr): continue
    out.append((r[hdr.index('Kernel Name')].split('(')[0][:50], r[hdr.index('Metric Value')], r[hdr.index('Metric Unit')]))
for k in out[-16:]: print(k)
PY
