#!/bin/bash
# weight-gradient GEMM of 1x1 layers: X operand tiled vs im2col TMA, column-tile width (boxes per tile) sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { VTB_GRAPH=1 timeout 120 tools/bench_conv $1 10 2>&1 | grep -E "graph replay" | grep "wgrad" | sed 's/(host-free, back to back)//g' | tr '\n' ' '; }
{
echo "--- X operand: tiled (default) vs im2col requests (VTB_WG_XTILED=0), default tile width"
for g in "256 22 22 128 128 1 1 0" "256 11 11 256 256 1 1 0" "256 88 88 64 64 1 1 0" "256 44 44 128 128 1 1 0" "256 22 22 256 256 1 1 0" "128 28 28 1312 512 1 1 0" "128 14 14 1728 768 1 1 0" "128 7 7 2144 1024 1 1 0" "128 56 56 768 256 1 1 0"; do
  for xt in 1 0; do VTB_WG_XTILED=$xt run "$g"; echo " <- xtiled=$xt $g"; done
done
echo "--- tile width sweep (boxes per tile), tiled X"
for g in "128 28 28 1312 512 1 1 0" "128 14 14 1728 768 1 1 0" "128 7 7 2144 1024 1 1 0" "128 56 56 768 256 1 1 0"; do
  for b in 0 2 3 4 6 7 8; do VTB_WG_BOXES=$b run "$g"; echo " <- boxes=$b $g"; done
done
} 2>&1 | tee gpurun_out/r02_wgrad_1x1_xtiled_sweep.txt
