#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_quick.json | cut -c1-260; tail -3 gpurun_out/bench.err
echo "=== layers"; timeout 300 python tools/layer_profile.py cspdarknet53 > gpurun_out/r02_layers_quick.txt 2>&1; head -${LAYER_LINES:-60} gpurun_out/r02_layers_quick.txt
echo done
