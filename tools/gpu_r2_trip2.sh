#!/bin/bash
# Round-2 visit 2: BatchNorm-backward statistics in the dgrad epilogue (vtb_conv_dgrad_bn): parity suite, A/B bench, layers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu_t2.log | tail -12
echo "=== bench (dgrad_bn on)"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t2.json | cut -c1-300; tail -3 gpurun_out/bench.err
echo "=== bench (dgrad_bn off)"; VTB_DGRAD_BN=0 timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t2_off.json | cut -c1-300
echo "=== layers"; timeout 300 python tools/layer_profile.py cspdarknet53 > gpurun_out/r02_layers_t2.txt 2>&1; head -50 gpurun_out/r02_layers_t2.txt
echo done
