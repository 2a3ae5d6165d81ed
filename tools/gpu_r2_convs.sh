#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for bn in 0 128 64; do
for g in "256 11 11 256 256 1 1 0" "256 11 11 256 256 3 1 1" "256 6 6 512 512 1 1 0" "256 6 6 512 512 3 1 1" "256 22 22 128 128 1 1 0"; do
  echo "--- VTB_BLOCK_N=$bn  $g"
  VTB_BLOCK_N=$bn VTB_GRAPH=1 timeout 120 tools/bench_conv $g 20 2>&1 | grep -E "graph replay|span" | grep -v "wgrad\|dgradbn\|dgr" | head -4
done; done 2>&1 | tee gpurun_out/r02_convs_blockn.txt
