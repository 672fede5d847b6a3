#!/bin/bash
# round 2, call Z: parity tests, then the default bench (all extras, configs 3-5 at full size)
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log | cut -c1-300
S=$(date +%s)
timeout -s KILL 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan frac', round(d['roofline']['frac'],3), 'pipe', round(d['roofline']['pipeline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})[:500])
except Exception as e:
    print('bench parse failed', e)
PY
