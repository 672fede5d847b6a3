#!/bin/bash
# round 2, call W (N GPUs): shard tests on one GPU, then the sharded headline, 3 runs
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 900 python -m pytest tests/test_shard.py tests/test_synth.py -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for rep in 1 2 3; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 5 --no-configs --no-extras --no-cpu > gpurun_out/bench_n${N}_w.log 2> gpurun_out/bench_n${N}_w.err; rc=$?
  grep -v "OMP_NUM\|^\*\*\*\|^$\|Warning\|NCCL version" gpurun_out/bench_n${N}_w.err | tail -5
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n${N}_w.log').read().strip().splitlines()[-1])
    print('rc $rc n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('rows_verified'), d['run']['sharded_rows_verified'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench parse failed', e)
PY
done
