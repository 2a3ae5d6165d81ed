#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== test_igemm"; bash tools/gpu_check_igemm.sh 2>&1 | grep -E "PASS|FAIL|exit code|bad [1-9]" | sort | uniq -c | head
for g in "256 176 176 32 32 1 1 0" "256 88 88 64 64 1 1 0" "256 22 22 128 128 1 1 0" "256 11 11 256 256 1 1 0" "256 22 22 128 128 3 1 1" "256 88 88 32 32 3 1 1"; do
  VTB_GRAPH=1 timeout 120 tools/bench_conv $g 10 2>&1 | grep -E "graph replay" | grep -v "wgrad" | tr '\n' ' '; echo " <- $g"
done 2>&1 | tee gpurun_out/r02_convs_epiopt.txt
echo "=== pytest (parity subset)"; timeout 900 python -m pytest tests/test_gpu_benchscale.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_quick.json | cut -c1-260
