#!/bin/bash
# ncu --set full of one general-path kernel ($K regex) from tools/general_prof.py
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-1} -c 1 -o gpurun_out/prof_gen python tools/general_prof.py > gpurun_out/ncu_gen.log 2>&1
tail -3 gpurun_out/ncu_gen.log
