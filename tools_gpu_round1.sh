#!/bin/bash
# First GPU pass: smoke, parity tests, short bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/pytest.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
