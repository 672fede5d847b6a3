#!/bin/bash
# round 2, final evidence pass of the build with the warp-per-chunk speculative pass and the bit-parallel FASTA kernels:
# smoke(), parity tests, default bench + reference arm, launch list and full ncu metrics of every device path
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
bash tools/gpu_r2_e.sh
