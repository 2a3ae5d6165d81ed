#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 0 1; do
for g in "256 176 176 32 32 1 1 0" "256 88 88 64 64 1 1 0" "256 88 88 32 32 3 1 1" "256 44 44 64 64 3 1 1"; do
  echo "--- VTB_STATS_DEFER=$d  $g"
  VTB_STATS_DEFER=$d VTB_GRAPH=1 timeout 120 tools/bench_conv $g 10 2>&1 | grep -E "^fprop" | head -2
done; done 2>&1 | tee gpurun_out/r02_convs_defer.txt
echo "=== pytest (parity subset)"; timeout 900 python -m pytest tests/test_gpu_benchscale.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_quick.json | cut -c1-260
echo "=== bench defer off"; VTB_STATS_DEFER=0 timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | cut -c1-260
