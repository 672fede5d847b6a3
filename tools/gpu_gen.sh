#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_general.csv python tools/general_prof.py > gpurun_out/gen.log 2>&1
tail -3 gpurun_out/gen.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_general.csv')) if len(r)>5]
hdr=next(r for r in rows if 'Kernel Name' in r)
last={}
order=[]
for r in rows:
    if r is hdr or len(r)!=len(hdr): continue
    k=r[hdr.index('Kernel Name')].split('(')[0][:50]
    v=float(r[hdr.index('Metric Value')].replace(',',''))
    u=r[hdr.index('Metric Unit')]
    v*= {'ns':1e-3,'us':1.0,'ms':1e3,'nsecond':1e-3,'usecond':1.0,'msecond':1e3}.get(u,1.0)
    order.append((k,v))
for k,v in order[-16:]: print('%-52s %10.1f us'%(k,v))
PY
