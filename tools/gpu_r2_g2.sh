#!/bin/bash
# round 2, call G2: parity tests + fuzz, launch list of the consumers (gather / sums / pack2)
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log | cut -c1-300
FUZZ_SECONDS=${FUZZ_SECONDS:-30} timeout -s KILL 400 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log | cut -c1-600
timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_g2.csv python tools/prof_paths.py ${PATHS:-gather sums pack2} > gpurun_out/prof_g2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_g2.csv')) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r)
for r in rows:
    if r is hdr or len(r)!=len(hdr): continue
    print('%-40s %-28s %s %s'%(r[hdr.index('Kernel Name')].split('(')[0][:40], r[hdr.index('Metric Name')], r[hdr.index('Metric Value')], r[hdr.index('Metric Unit')]))
PY
