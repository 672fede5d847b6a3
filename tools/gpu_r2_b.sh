#!/bin/bash
# round 2, call B: new tests + bench configs only
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_synth.py tests/test_shard.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "synth or sharded_parse_of or single_shard or arrayadd or stateless or device_generator" > gpurun_out/pytest_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_b.log
tail -15 gpurun_out/pytest_b.log
S=$(date +%s)
timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
tail -c 1500 gpurun_out/bench_b.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_b.log').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})[:900])
except Exception as e:
    print('bench parse failed', e)
PY
