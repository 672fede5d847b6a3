#!/bin/bash
# round 2: the default bench exactly as the driver launches it at N ranks (and the reference arm on rank 0)
mkdir -p gpurun_out
N=${NGPU:-4}
S=$(date +%s)
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps ${STEPS:-200} --warmup 5 > gpurun_out/bench_full_n$N.log 2> gpurun_out/bench_full_n$N.err; echo "full bench N=$N exit $? in $(( $(date +%s) - S )) s"
grep -v "OMP_NUM\|^\*\*\*\|^$\|Warning\|NCCL version" gpurun_out/bench_full_n$N.err | tail -8
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_full_n$N.log').read().strip().splitlines()[-1])
    print('n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'rows', d['run']['sharded_rows_verified'], 'e2e', json.dumps(d['e2e'])[:400])
    for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})[:500])
except Exception as e:
    print('bench parse failed', e)
PY
