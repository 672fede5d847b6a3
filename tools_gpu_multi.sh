#!/bin/bash
# multi-GPU pass: shard tests on one GPU, then bench at N GPUs under torchrun
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 900 python -m pytest tests/test_shard.py tests/test_gpu_parity.py -q -m gpu -x --timeout 600 > gpurun_out/pytest_shard.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_shard.log
tail -15 gpurun_out/pytest_shard.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -3 gpurun_out/bench_n$N.log; tail -20 gpurun_out/bench_n$N.err
