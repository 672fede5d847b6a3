#!/bin/bash
# round 2, call FB: parity tests + fuzz, FASTA step time and the launch list of its kernels
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log | cut -c1-300
FUZZ_SECONDS=${FUZZ_SECONDS:-40} FUZZ_KINDS=${FUZZ_KINDS:-} timeout -s KILL 400 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log | cut -c1-600
bash tools/gpu_r2_fa.sh
