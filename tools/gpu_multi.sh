#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 900 python -m pytest tests/test_shard.py -q -m gpu -x --timeout 600 > gpurun_out/pytest_shard.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_shard.log
tail -5 gpurun_out/pytest_shard.log
for T in ${TRANSPORTS:-fused peer nccl}; do
FQB_SHARD_TRANSPORT=$T timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 --e2e-steps 2 > gpurun_out/bench_n${N}_$T.log 2> gpurun_out/bench_n${N}_$T.err; echo "bench N=$N $T exit $?"
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n${N}_$T.err | tail -8
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$T.log').read().strip().splitlines()[-1])
    print('$T', 'n', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), d['config']['sharding'], 'records', d['config']['records_per_gpu'])
except Exception as e: print('failed', e)
PY
done
