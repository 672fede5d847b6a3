#!/bin/bash
# round 2, call T (N GPUs): the shard epilogue in the scan (FQB_SHARD_TAIL) against the two small kernels, 3 runs each
mkdir -p gpurun_out
N=${NGPU:-2}
python tools/ab_paths.py fast 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for rep in 1 2 3; do for tl in 0 1; do
  FQB_SHARD_TAIL=$tl timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 5 --no-configs --no-extras --no-cpu > gpurun_out/bench_n${N}_t$tl.log 2> gpurun_out/bench_n${N}_t$tl.err; rc=$?
  grep -v "OMP_NUM\|^\*\*\*\|^$\|Warning\|NCCL version" gpurun_out/bench_n${N}_t$tl.err | tail -5
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n${N}_t$tl.log').read().strip().splitlines()[-1])
    print('rc $rc tail=$tl n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('rows_verified'), d['run']['sharded_rows_verified'])
except Exception as e:
    print('bench parse failed', e)
PY
done; done
