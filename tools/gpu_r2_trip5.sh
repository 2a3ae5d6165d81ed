#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== launch overhead"; timeout 120 tools/launch_overhead | tee gpurun_out/r02_launch_overhead.txt
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_t5.log
echo "=== bench default"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t5.json | cut -c1-260; tail -3 gpurun_out/bench.err
echo done
