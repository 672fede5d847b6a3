#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 100 --no-cpu > gpurun_out/bench_extras.log 2> gpurun_out/bench_extras.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_extras.log').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4)); print(json.dumps(d['extras'], indent=1))
PY
tail -3 gpurun_out/bench_extras.err
