#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_t9.log
echo "=== bench default (native SGD)"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t9.json | cut -c1-260; tail -3 gpurun_out/bench.err
echo "=== bench torch SGD"; VTB_NATIVE_SGD=0 timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t9_torchsgd.json | cut -c1-260
echo "=== bench native head too"; VTB_NATIVE_HEAD=1 timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t9_nativehead.json | cut -c1-260
echo done
