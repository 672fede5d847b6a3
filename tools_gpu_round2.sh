#!/bin/bash
# GPU pass: parity tests, bench for each scan configuration, ncu launch list + full capture.
mkdir -p gpurun_out
ls -la oracle/_ref/fastqandfurious/ > gpurun_out/ref_ls.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
for c in 0 1 2 3; do
  timeout -s KILL 600 python bench.py --steps 100 --warmup 3 --cfg $c --no-cpu --no-extras > gpurun_out/bench_cfg$c.log 2> gpurun_out/bench_cfg$c.err; echo "cfg $c exit $?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_cfg$c.log').read().strip().splitlines()[-1])
    print('cfg', $c, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel_config'])
except Exception as e:
    print('cfg $c parse failed', e)
PY
done
timeout -s KILL 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -s 3 -c 2 -o gpurun_out/prof_scan python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_scan.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fq_emit_kernel -s 3 -c 2 -o gpurun_out/prof_emit python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_emit.log 2>&1
ls -la gpurun_out
