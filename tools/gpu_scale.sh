#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout -s KILL 600 python bench.py --steps 100 --no-cpu > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "N=1 exit $?"
for N in ${NS:-2 4 8}; do
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'n', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'extras', {k: round(v.get('gbs', v.get('hbm_gbs', 0)),1) for k,v in (d.get('extras') or {}).items()})
    except Exception as e:
        print(f, 'failed', e)
PY
