#!/bin/bash
# round 2, call I (2 GPUs): sharded headline with and without double-buffered pull-ahead; source-level ncu of the new gspec
mkdir -p gpurun_out
N=${NGPU:-2}
for dbl in 0 1 0 1; do
  FQB_SHARD_DOUBLE=$dbl timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 --no-configs --no-extras --no-cpu > gpurun_out/bench_n${N}_dbl$dbl.log 2> gpurun_out/bench_n${N}_dbl$dbl.err; echo "bench N=$N double=$dbl exit $?"
  grep -v "OMP_NUM\|^\*\*\*\|^$\|Warning" gpurun_out/bench_n${N}_dbl$dbl.err | tail -5
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n${N}_dbl$dbl.log').read().strip().splitlines()[-1])
    print('double=$dbl n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['run']['sharded_rows_verified'])
except Exception as e:
    print('bench parse failed', e)
PY
done
k=fq_gspec_kernel; pth=multiline_spec
timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/src_${k}_v2 python tools/prof_paths.py $pth > gpurun_out/ncu_src_${k}.log 2>&1; echo "ncu $k exit $?"
ncu -i gpurun_out/src_${k}_v2.ncu-rep --page source --csv > gpurun_out/src_${k}_v2.csv 2>/dev/null
ncu -i gpurun_out/src_${k}_v2.ncu-rep --page raw --csv > gpurun_out/raw_${k}_v2.csv 2>/dev/null
rm -f gpurun_out/src_${k}_v2.ncu-rep
