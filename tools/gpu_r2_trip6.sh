#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== launch overhead"; timeout 120 tools/launch_overhead | tee gpurun_out/r02_launch_overhead.txt
echo "=== bench default"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t6.json | cut -c1-260; tail -3 gpurun_out/bench.err
echo "=== layers"; timeout 300 python tools/layer_profile.py cspdarknet53 > gpurun_out/r02_layers_t6.txt 2>&1; grep -E "total|dgrad|host" gpurun_out/r02_layers_t6.txt | head -40
echo done
