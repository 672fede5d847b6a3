#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/shard_timeline.py 2>&1 | grep -E "^rank|Error|error" | sort | tee gpurun_out/shard_timeline_n$N.log
