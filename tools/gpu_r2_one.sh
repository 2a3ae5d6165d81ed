#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest parity subset with VTB_FUSED_NORM=1"; VTB_FUSED_NORM=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchscale.py -m gpu -q 2>&1 | tail -4
echo "=== bench VTB_FUSED_NORM=1"; VTB_FUSED_NORM=1 timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t16_fusednorm.json | cut -c1-200; tail -2 gpurun_out/bench.err
echo "=== bench default"; timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t16.json | cut -c1-200; tail -2 gpurun_out/bench.err
echo "=== layers VTB_FUSED_NORM=1"; VTB_FUSED_NORM=1 timeout 300 python tools/layer_profile.py cspdarknet53 > gpurun_out/r02_layers_fusednorm.txt 2>&1; grep -E "fprop|bn_act" gpurun_out/r02_layers_fusednorm.txt | head -24 | cut -c1-120
