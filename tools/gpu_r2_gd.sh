#!/bin/bash
# round 2: parity tests + fuzz of the build (FASTA at 32 KiB per iteration, geometry from the line density), step times
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log | cut -c1-250
FUZZ_SECONDS=${FUZZ_SECONDS:-40} timeout -s KILL 400 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -2 gpurun_out/fuzz.log | cut -c1-600
timeout -s KILL 300 python tools/ab_paths.py multiline fasta 2>&1 | grep -v Warning | tee gpurun_out/ab_gd.log
