#!/bin/bash
# development visit: kernel harness + wgrad timing + parity tests + per-layer profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_check_igemm.sh > /dev/null 2>&1; grep -E "FAIL|exit code|error|mismatch|timeout" gpurun_out/igemm_check.log | sort | uniq -c | sort -rn | head -20; grep -c " ok " gpurun_out/igemm_check.log
B=tools/bench_conv
run() { echo "## $*"; $B "$@" | grep -A1 wgrad; }
run 256 22 22 128 128 3 1 1
run 256 6 6 512 512 3 1 1
run 256 22 22 128 128 1 1 0
run 256 88 88 64 32 1 1 0
run 256 11 11 512 1024 3 2 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/layer_profile.py ${1:-cspdarknet53} > gpurun_out/layers.txt 2>&1; head -${2:-60} gpurun_out/layers.txt
