#!/bin/bash
# differential fuzzing of the CUDA paths against the oracle (tests/fuzz_gpu.py); FUZZ_SECONDS / FUZZ_SEED
mkdir -p gpurun_out
timeout -s KILL $(( ${FUZZ_SECONDS:-60} + 120 )) python tests/fuzz_gpu.py > gpurun_out/fuzz.log 2>&1; echo "fuzz exit $?"
tail -12 gpurun_out/fuzz.log
