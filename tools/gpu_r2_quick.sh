#!/bin/bash
# quick visit: benchscale parity + unit tests, bench, per-layer table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest (parity subset)"; timeout 900 python -m pytest tests/test_gpu_benchscale.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_quick.json | cut -c1-260; tail -3 gpurun_out/bench.err
echo "=== layers"; timeout 300 python tools/layer_profile.py cspdarknet53 > gpurun_out/r02_layers_quick.txt 2>&1; grep -E "total|wgrad" gpurun_out/r02_layers_quick.txt | head -${LAYER_LINES:-14}
echo done
