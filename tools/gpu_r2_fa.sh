#!/bin/bash
# round 2, call FA: FASTA step time and the launch list of its kernels
mkdir -p gpurun_out
timeout -s KILL 300 python tools/ab_paths.py fasta 2>&1 | grep -v Warning | tee gpurun_out/ab_fa.log
timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/launches_fa.csv python tools/prof_paths.py fasta > gpurun_out/prof_fa.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_fa.csv')) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r)
for r in rows:
    if r is hdr or len(r)!=len(hdr): continue
    print('%-40s %-28s %s %s'%(r[hdr.index('Kernel Name')].split('(')[0][:40], r[hdr.index('Metric Name')], r[hdr.index('Metric Value')], r[hdr.index('Metric Unit')]))
PY
