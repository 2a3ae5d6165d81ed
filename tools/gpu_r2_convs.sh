#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in "256 22 22 128 128 1 1 0" "256 11 11 256 256 1 1 0" "256 6 6 512 512 3 1 1"; do
  VTB_GRAPH=1 timeout 120 tools/bench_conv $g 20
done 2>&1 | grep -v "dgradbn\|dgr+acc" | tee gpurun_out/r02_convs2.txt
