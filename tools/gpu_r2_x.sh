#!/bin/bash
# round 2, call X: the warp-per-chunk speculative pass (fq_gspec2.cuh) -- parity tests, fuzz, A/B against the CTA-per-chunk kernel
mkdir -p gpurun_out
S=$(date +%s)
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log | cut -c1-400
for v in 1 2; do
  FQB200_SPEC=$v timeout -s KILL 300 python tools/ab_paths.py multiline 2>&1 | grep -v Warning | sed "s/^/spec v$v: /" | tee -a gpurun_out/ab_x.log
done
FUZZ_SECONDS=${FUZZ_SECONDS:-60} timeout -s KILL 400 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log | cut -c1-600
