#!/bin/bash
# weight-gradient GEMM: column-tile width (boxes per tile) sweep on the 1x1 layers (split count / partial traffic trade-off)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in "256 22 22 128 128 1 1 0" "256 11 11 256 256 1 1 0" "256 6 6 512 512 1 1 0" "256 44 44 64 64 1 1 0" "256 6 6 1024 1024 1 1 0" "256 44 44 128 128 1 1 0" "256 22 22 256 256 1 1 0" "256 11 11 512 512 1 1 0" "256 88 88 64 64 1 1 0"; do
  for b in 0 1 2; do
    VTB_WG_BOXES=$b VTB_GRAPH=1 timeout 120 tools/bench_conv $g 10 2>&1 | grep -E "graph replay" | grep "wgrad" | sed 's/(host-free, back to back)//g' | tr '\n' ' '; echo " <- boxes=$b $g"
  done
done 2>&1 | tee gpurun_out/r02_wgrad_boxes_sweep.txt
