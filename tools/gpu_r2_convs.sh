#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ks in 0 1; do
for g in "256 176 176 32 32 1 1 0" "256 88 88 64 64 1 1 0" "256 88 88 64 32 1 1 0" "256 88 88 32 32 3 1 1" "256 44 44 64 64 3 1 1" "256 44 44 128 64 1 1 0"; do
  VTB_KSPLIT=$ks VTB_GRAPH=1 timeout 120 tools/bench_conv $g 10 2>&1 | grep -E "graph replay" | grep -v "wgrad\|dgradbn" | sed 's/(host-free, back to back)//g' | tr '\n' ' '; echo " <- ks=$ks $g"
done; done 2>&1 | tee gpurun_out/r02_convs_ksplit.txt
