#!/bin/bash
# round 2, call K (2 GPUs): parity tests, multiline timing, sharded headline single / double buffered (3 runs each),
# then the full default bench at N ranks (configs, extras, sharded end-to-end)
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python tools/ab_paths.py multiline 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for rep in 1 2 3; do for dbl in 0 1; do
  FQB_SHARD_DOUBLE=$dbl timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 5 --no-configs --no-extras --no-cpu > gpurun_out/bench_n${N}_d$dbl.log 2> gpurun_out/bench_n${N}_d$dbl.err; rc=$?
  grep -v "OMP_NUM\|^\*\*\*\|^$\|Warning\|NCCL version" gpurun_out/bench_n${N}_d$dbl.err | tail -5
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n${N}_d$dbl.log').read().strip().splitlines()[-1])
    print('rc $rc double=$dbl n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('sharded'), d['e2e'].get('rows_verified'), d['run']['sharded_rows_verified'])
except Exception as e:
    print('bench parse failed', e)
PY
done; done
S=$(date +%s)
FQB_SHARD_DOUBLE=${FULL_DOUBLE:-1} timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N > gpurun_out/bench_full_n$N.log 2> gpurun_out/bench_full_n$N.err; echo "full bench N=$N exit $? in $(( $(date +%s) - S )) s"
grep -v "OMP_NUM\|^\*\*\*\|^$\|Warning\|NCCL version" gpurun_out/bench_full_n$N.err | tail -8
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_full_n$N.log').read().strip().splitlines()[-1])
    print('n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', json.dumps(d['e2e'])[:600])
    for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})[:700])
except Exception as e:
    print('bench parse failed', e)
PY
