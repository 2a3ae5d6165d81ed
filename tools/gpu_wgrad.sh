#!/bin/bash
# wgrad development visit: correctness harness + timing sweep of the wgrad kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != "nocheck" ]; then
bash tools/gpu_check_igemm.sh > /dev/null 2>&1; grep -E "FAIL|exit code|error|mismatch|timeout" gpurun_out/igemm_check.log | sort | uniq -c | sort -rn | head -20; grep -c " ok " gpurun_out/igemm_check.log
fi
B=tools/bench_conv
run() { echo "## $*"; $B "$@" | grep -A1 wgrad; echo -n "GEMM only: "; VTB_WG_NOREDUCE=1 $B "$@" | grep wgrad; }
run 256 22 22 128 128 3 1 1
echo -n "kpix128: "; VTB_WG_KPIX=128 VTB_WG_NOREDUCE=1 $B 256 22 22 128 128 3 1 1 | grep -A1 wgrad
echo -n "boxes2 : "; VTB_WG_BOXES=2 VTB_WG_NOREDUCE=1 $B 256 22 22 128 128 3 1 1 | grep -A1 wgrad
echo -n "boxes4 : "; VTB_WG_BOXES=4 VTB_WG_NOREDUCE=1 $B 256 22 22 128 128 3 1 1 | grep -A1 wgrad
run 256 11 11 256 256 3 1 1
run 256 6 6 512 512 3 1 1
run 256 176 176 16 32 3 1 1
run 256 176 176 32 64 3 2 1
run 256 88 88 32 32 3 1 1
run 256 44 44 64 64 3 1 1
run 256 88 88 64 32 1 1 0
run 256 22 22 128 128 1 1 0
run 256 11 11 512 1024 3 2 1
