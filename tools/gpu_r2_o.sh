#!/bin/bash
# round 2, call O: parity + fuzz, multiline timing, racecheck, gspec launch time
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
FUZZ_SECONDS=${FUZZ_SECONDS:-60} timeout -s KILL 600 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log
python tools/ab_paths.py ${AB_PATHS:-multiline illumina} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
TOOLS=racecheck bash tools/gpu_sanitize.sh
tail -3 gpurun_out/sanitize_racecheck.log
timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/launches_o.csv python tools/prof_paths.py multiline_spec > gpurun_out/prof_o.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_o.csv')) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r)
for r in rows:
    if r is hdr or len(r)!=len(hdr): continue
    print('%-40s %-28s %s %s'%(r[hdr.index('Kernel Name')].split('(')[0][:40], r[hdr.index('Metric Name')], r[hdr.index('Metric Value')], r[hdr.index('Metric Unit')]))
PY
