#!/bin/bash
# tests + default bench with extras (+ general-path launch list when GEN=1)
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log
timeout -s KILL 900 python bench.py --steps 100 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    for k, v in (d.get('extras') or {}).items(): print(' ', k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
if [ -n "$GEN" ]; then bash tools/gpu_gen.sh; fi
