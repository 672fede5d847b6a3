#!/bin/bash
# round 2, call H: parity + fuzz of the current build, A/B step times, launch list of the general path
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
FUZZ_SECONDS=${FUZZ_SECONDS:-90} timeout -s KILL 600 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log
python tools/ab_paths.py ${AB_PATHS} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for v in ${VARIANTS}; do
  FQB200_LIB=$PWD/$v python tools/ab_paths.py ${AB_PATHS} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
done
timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_h.csv python tools/prof_paths.py ${PROF_PATHS:-fast ont multiline_spec} > gpurun_out/prof_h.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_h.csv')) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r)
for r in rows:
    if r is hdr or len(r)!=len(hdr): continue
    print('%-50s %s %s'%(r[hdr.index('Kernel Name')].split('(')[0][:50], r[hdr.index('Metric Value')], r[hdr.index('Metric Unit')]))
PY
if [ -n "$BENCH" ]; then
timeout -s KILL 900 python bench.py --steps 100 --no-configs --no-cpu --no-extras > gpurun_out/bench_h.log 2> gpurun_out/bench_h.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_h.log').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan frac', round(d['roofline']['frac'],3), 'pipe', round(d['roofline']['pipeline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'with events', d['roofline']['ms_per_step_with_kernel_events'])
PY
fi
