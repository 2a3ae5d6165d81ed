#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in new r1 new r1; do echo "=== C5 rule=$r"; VTB_BLOCKM_RULE=$r timeout 300 python bench.py --config C5 --no-cpu-baseline 2>/dev/null | cut -c1-150; done
for r in new r1; do echo "=== C2 rule=$r"; VTB_BLOCKM_RULE=$r timeout 300 python bench.py --config C2 --no-cpu-baseline 2>/dev/null | cut -c1-150; done
