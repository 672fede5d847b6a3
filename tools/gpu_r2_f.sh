#!/bin/bash
# round 2, call F: parity tests + fuzz of the current build, then A/B step times of the library variants under variants/
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
FUZZ_SECONDS=${FUZZ_SECONDS:-60} timeout -s KILL 600 python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/fuzz.log
for rep in 1 2; do
python tools/ab_paths.py ${AB_PATHS} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
for v in variants/*.so; do
  FQB200_LIB=$PWD/$v python tools/ab_paths.py ${AB_PATHS} 2>&1 | grep -v Warning | tee -a gpurun_out/ab.log
done
done
