#!/bin/bash
# round 2, final evidence pass: smoke(), parity tests, default bench + reference arm, launch list of every path
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
bash tools/gpu_r2_e.sh
