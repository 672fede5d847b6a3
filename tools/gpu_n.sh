#!/bin/bash
# bench at N GPUs (default transport), short
mkdir -p gpurun_out
N=${NGPU:-8}
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 --e2e-steps 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -6
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print('n', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['config']['sharding'], 'records', d['config']['records_per_gpu'], d.get('extras'))
except Exception as e: print('failed', e)
PY
