#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -6 gpurun_out/pytest.log
./tools_gpu_gen.sh
