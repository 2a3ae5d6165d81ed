#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in "256 22 22 128 128 1 1 0" "256 22 22 128 128 3 1 1" "256 88 88 64 64 1 1 0" "256 11 11 256 256 1 1 0"; do
  VTB_GRAPH=1 timeout 120 tools/bench_conv $g 20
done 2>&1 | tee gpurun_out/r02_convs.txt
