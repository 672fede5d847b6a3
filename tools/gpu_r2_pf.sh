#!/bin/bash
# round 2, experiment: L2 prefetch of later tiles in the scan kernel (variants/libfqb200_pfN.so against the stock build)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "speculative or config_shapes" 2>&1 | tail -2
for rep in 1 2; do
for lib in "" variants/libfqb200_pf2.so variants/libfqb200_pf4.so; do
  FQB200_LIB=${lib:+$PWD/$lib} timeout -s KILL 300 python tools/ab_paths.py fast ont multiline 2>&1 | grep -v "Warning\|rows verified" | tee -a gpurun_out/ab_pf.log
done; done
