mkdir -p gpurun_out
for v in stock FQB_SCAN_FMA_ADDS FQB_SCAN_FMA_HALF stock2 FQB_SCAN_FMA_ADDS2; do
  lib=""
  case $v in FQB_SCAN_FMA_ADDS*) lib=$PWD/tools/_san/libfqb200_FQB_SCAN_FMA_ADDS.so;; FQB_SCAN_FMA_HALF) lib=$PWD/tools/_san/libfqb200_FQB_SCAN_FMA_HALF.so;; esac
  FQB200_LIB=$lib timeout -s KILL 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1])
print('$v', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'scan_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3))
PY
done
FQB200_LIB=$PWD/tools/_san/libfqb200_FQB_SCAN_FMA_ADDS.so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fixed150 or clean_input or corpus_auto" 2>&1 | tail -3
