#!/bin/bash
# round 2, call A: parity tests, the default bench (configs 3-5 included), launch list + full ncu of the general path
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
S=$(date +%s)
timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    print('cpu', json.dumps(d.get('cpu_baseline'))[:1500])
    print('dropin', d['e2e'].get('dropin_per_record'))
    for k, v in (d.get('extras') or {}).items(): print(' ', k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})[:600])
except Exception as e:
    print('bench parse failed', e)
PY
bash tools/gpu_gen.sh
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fq_g_ -c 14 -o gpurun_out/prof_general python tools/general_prof.py > gpurun_out/ncu_general.log 2>&1; echo "ncu general exit $?"
