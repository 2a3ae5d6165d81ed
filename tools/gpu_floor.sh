#!/bin/bash
cd "$(dirname "$0")/.."
B=tools/bench_conv
export VTB_GRAPH=1
for pdl in 1 0; do
echo "#### VTB_PDL=$pdl"
export VTB_PDL=$pdl
echo "## tiny"; $B 1 6 6 64 64 1 1 0 | grep -E "graph"
echo "## 256 6 6 512 512 1x1"; $B 256 6 6 512 512 1 1 0 | grep -E "graph| us "
echo "## 256 11 11 256 256 1x1"; $B 256 11 11 256 256 1 1 0 | grep -E "graph| us "
echo "## 256 22 22 128 128 1x1"; $B 256 22 22 128 128 1 1 0 | grep -E "graph| us "
echo "## 256 22 22 128 128 3x3"; $B 256 22 22 128 128 3 1 1 | grep -E "graph| us "
echo "## 256 44 44 64 64 1x1"; $B 256 44 44 64 64 1 1 0 | grep -E "graph| us "
done
