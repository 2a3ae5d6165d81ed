#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== memcheck: head kernels"; timeout 900 compute-sanitizer --tool memcheck --kernel-name regex:"tile_gemm|head_" python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "native_head" 2>&1 | tail -8 | tee gpurun_out/r02_sanitizer_memcheck_head.log
echo "=== racecheck: head kernels"; timeout 900 compute-sanitizer --tool racecheck --kernel-name regex:"tile_gemm|head_" python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "native_head" 2>&1 | tail -8 | tee gpurun_out/r02_sanitizer_racecheck_head.log
echo done
