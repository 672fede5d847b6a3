#!/bin/bash
# round 2, experiment: tiles per chunk of the warp-per-chunk pass (window fill 70 / 85 / 100 % -> 3 / 4 / 5 tiles at 22 M lines per GiB)
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "" variants/libfqb200_fill70.so variants/libfqb200_fill100.so; do
  FQB200_LIB=${lib:+$PWD/$lib} timeout -s KILL 300 python tools/ab_paths.py multiline 2>&1 | grep -v "Warning" | tee -a gpurun_out/ab_fill.log
done; done
