#!/bin/bash
# compute-sanitizer over tests/sanitize_workload.py: memcheck, racecheck (shared memory), synccheck.
# racecheck runs on a build with -DFQB_NO_DUMMY_STORE (tools/_san/libfqb200_nodummy.so, built by
#   nvcc <flags of __graft_entry__.py> -DFQB_NO_DUMMY_STORE csrc/fqb200.cu): the scan kernel's branch-free queue
# store sends the lanes WITHOUT a newline to one never-read dummy word per warp, which racecheck (rightly) reports
# as write-write overlap; the variant predicates that store instead, everything else is identical.
mkdir -p gpurun_out tools/_san
nodummy=tools/_san/libfqb200_nodummy.so
if [ ! -f $nodummy ] || [ -n "$(find fastq-and-furious_b200/csrc include -newer $nodummy -type f | head -1)" ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DFQB_NO_DUMMY_STORE \
    fastq-and-furious_b200/csrc/fqb200.cu -o $nodummy 2> /dev/null || echo "could not build $nodummy"
fi
timeout -s KILL 200 python tests/sanitize_workload.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit $?"
tail -2 gpurun_out/sanitize_plain.log
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  lib=""
  if [ $tool = racecheck ] && [ -z "$RACE_STOCK" ]; then lib=$PWD/tools/_san/libfqb200_nodummy.so; fi
  FQB200_LIB=$lib timeout -s KILL ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 40 python tests/sanitize_workload.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $? (lib: ${lib:-stock})"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error:|Hazard|hazard" gpurun_out/sanitize_$tool.log | head -12
done
