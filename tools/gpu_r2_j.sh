#!/bin/bash
# round 2, call J (2 GPUs): parity + fuzz of the current build, multiline A/B, sharded headline single / double buffered
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
FUZZ_SECONDS=${FUZZ_SECONDS:-60} timeout -s KILL 600 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log
python tools/ab_paths.py multiline illumina 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for mode in "0 0" "1 0" "1 1" "0 0" "1 0"; do
  set -- $mode
  FQB_SHARD_DOUBLE=$1 FQB_SHARD_TAIL=$2 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 --no-configs --no-extras --no-cpu > gpurun_out/bench_n${N}_d$1t$2.log 2> gpurun_out/bench_n${N}_d$1t$2.err; echo "bench N=$N double=$1 tail=$2 exit $?"
  grep -v "OMP_NUM\|^\*\*\*\|^$\|Warning\|NCCL version" gpurun_out/bench_n${N}_d$1t$2.err | tail -5
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n${N}_d$1t$2.log').read().strip().splitlines()[-1])
    print('double=$1 tail=$2 n', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['run']['sharded_rows_verified'])
except Exception as e:
    print('bench parse failed', e)
PY
done
k=fq_gspec_kernel; pth=multiline_spec
timeout -s KILL 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/src_${k}_v3 python tools/prof_paths.py $pth > gpurun_out/ncu_src_${k}.log 2>&1; echo "ncu $k exit $?"
ncu -i gpurun_out/src_${k}_v3.ncu-rep --page source --csv > gpurun_out/src_${k}_v3.csv 2>/dev/null
ncu -i gpurun_out/src_${k}_v3.ncu-rep --page raw --csv > gpurun_out/raw_${k}_v3.csv 2>/dev/null
rm -f gpurun_out/src_${k}_v3.ncu-rep
