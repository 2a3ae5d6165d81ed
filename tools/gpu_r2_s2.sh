#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_final.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke_final.log
echo "=== bench"; timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_final.json | cut -c1-200; tail -2 gpurun_out/bench.err
