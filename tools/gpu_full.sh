#!/bin/bash
# full GPU pass: all gpu tests, default bench (with extras + cpu baseline), reference arm, ncu launch list + full captures
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $?"
cat gpurun_out/bench_ref.log | cut -c1-400
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fq_scan_kernel -s 3 -c 1 -o gpurun_out/prof_scan python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_scan.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fq_emit_kernel -s 3 -c 1 -o gpurun_out/prof_emit python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_emit.log 2>&1
