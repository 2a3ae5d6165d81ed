#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest head/loss"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "head or loss or training or benchscale" 2>&1 | tail -3
echo "=== bench default (native head)"; timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t12.json | cut -c1-200; tail -2 gpurun_out/bench.err
echo "=== bench VTB_NATIVE_HEAD=0"; VTB_NATIVE_HEAD=0 timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t12_torchhead.json | cut -c1-200
echo "=== wgrad 1x1 after heuristic"
for g in "256 11 11 256 256 1 1 0" "256 6 6 512 512 1 1 0" "256 22 22 256 256 1 1 0"; do
    VTB_GRAPH=1 timeout 120 tools/bench_conv $g 10 2>&1 | grep -E "graph replay" | grep "wgrad" | sed 's/(host-free, back to back)//g' | tr '\n' ' '; echo " <- $g"
done
echo done
