#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_r2_wgsweep.sh
echo "=== bench default"; timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t11.json | cut -c1-200
echo "=== bench VTB_BWD_COOP=1"; VTB_BWD_COOP=1 timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t11_coop.json | cut -c1-200
echo done
