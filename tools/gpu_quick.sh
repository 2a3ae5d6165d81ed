#!/bin/bash
# quick GPU visit: kernel check harness + parity tests + per-layer profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_check_igemm.sh > /dev/null 2>&1; grep -E "FAIL|PASS|exit code|error|mismatch" gpurun_out/igemm_check.log | sort | uniq -c | sort -rn | head -20
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/layer_profile.py ${1:-cspdarknet53} > gpurun_out/layers.txt 2>&1; head -70 gpurun_out/layers.txt
