#!/bin/bash
# Runs every case of tools/test_igemm in its own process (a trapped kernel poisons the CUDA context).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(tools/test_igemm count)
fail=0
for i in $(seq 0 $((n-1))); do
  timeout 120 tools/test_igemm $i 2>&1 | tail -40
  rc=${PIPESTATUS[0]}
  if [ $rc -ne 0 ]; then fail=$((fail+1)); echo "case $i exit code $rc"; fi
done 2>&1 | tee gpurun_out/igemm_check.log
echo "done"
