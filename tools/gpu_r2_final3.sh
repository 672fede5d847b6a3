#!/bin/bash
# round 2, last evidence pass: smoke(), parity tests, default bench + reference arm (no ncu)
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -2 gpurun_out/pytest.log | cut -c1-200
S=$(date +%s)
timeout -s KILL 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $? in $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan frac', round(d['roofline']['frac'],3), 'pipe', round(d['roofline']['pipeline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ('gbs','ms_per_step','mrec_s','hbm_gbs')}))
r = json.loads(open('gpurun_out/bench_ref.log').read().strip().splitlines()[-1]); print('reference', round(r['value'],2), r['unit'])
PY
