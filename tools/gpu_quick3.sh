#!/bin/bash
# A/B: weight-gradient GEMMs on a second stream
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu (VTB_WGRAD_STREAM=1)"; VTB_WGRAD_STREAM=1 timeout 900 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -30
echo "=== bench base"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench_base.json | cut -c1-200; tail -3 gpurun_out/bench.err
echo "=== bench VTB_WGRAD_STREAM=1"; VTB_WGRAD_STREAM=1 timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench_ws.err | tee gpurun_out/bench_ws.json | cut -c1-200; tail -3 gpurun_out/bench_ws.err
