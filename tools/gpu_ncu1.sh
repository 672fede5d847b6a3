#!/bin/bash
# one ncu --set full capture of kernel $K (regex) from the default bench (with extras)
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-2} -c 1 -o gpurun_out/prof_one python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_one.log 2>&1
tail -3 gpurun_out/ncu_one.log
