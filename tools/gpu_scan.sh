#!/bin/bash
# scan-kernel iteration: parity tests, bench per configuration (no cpu/extras), ncu of the scan kernel
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for c in ${CFGS:-0 1 2}; do
  timeout -s KILL 600 python bench.py --steps 100 --warmup 3 --cfg $c --no-cpu ${EXTRAS:---no-extras} > gpurun_out/bench_cfg$c.log 2> gpurun_out/bench_cfg$c.err; echo "cfg $c exit $?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_cfg$c.log').read().strip().splitlines()[-1])
    print('cfg', $c, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel_config'], d.get('extras'))
except Exception as e:
    print('cfg $c parse failed', e); print(open('gpurun_out/bench_cfg$c.err').read()[-2000:])
PY
done
if [ -n "$NCU" ]; then
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 --cfg ${PROF_CFG:-0} > gpurun_out/ncu_scan.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fq_emit_kernel -s 3 -c 1 -o gpurun_out/prof_emit python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 --cfg ${PROF_CFG:-0} > gpurun_out/ncu_emit.log 2>&1
fi
