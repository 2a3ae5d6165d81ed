#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in "256 11 11 256 256 1 1 0" "256 22 22 128 128 1 1 0" "256 11 11 256 256 3 1 1" "256 6 6 512 512 3 1 1"; do
  echo "--- $g"; VTB_INTERLEAVE=1 VTB_GRAPH=1 timeout 120 tools/bench_conv $g 10 2>&1 | grep -E "graph replay|interleave" | grep -v "dgradbn\|dgr+acc" | sed 's/(host-free, back to back)//g'
done 2>&1 | tee gpurun_out/r02_convs_interleave.txt
