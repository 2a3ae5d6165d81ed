#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_latest.log
echo "=== bench default (optimizer overlap)"; timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t14.json | cut -c1-200; tail -2 gpurun_out/bench.err
echo "=== bench VTB_SGD_OVERLAP=0"; VTB_SGD_OVERLAP=0 timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t14_nooverlap.json | cut -c1-200; tail -2 gpurun_out/bench.err
echo done
