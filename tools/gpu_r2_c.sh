#!/bin/bash
# round 2, call C (N GPUs): new tests on one GPU, then the bench at N ranks with the configs
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 900 python -m pytest tests/test_synth.py tests/test_shard.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "synth or sharded_parse_of or single_shard or arrayadd or stateless or device_generator" > gpurun_out/pytest_c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c.log
tail -5 gpurun_out/pytest_c.log
S=$(date +%s)
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_c_n$N.log 2> gpurun_out/bench_c_n$N.err; echo "bench N=$N exit $? in $(( $(date +%s) - S )) s"
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_c_n$N.err | tail -12
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c_n$N.log').read().strip().splitlines()[-1])
    print('n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['run'])
    for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})[:700])
except Exception as e:
    print('bench parse failed', e)
PY
