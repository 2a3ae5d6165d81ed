#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_check_igemm.sh > /dev/null 2>&1; grep -c " ok " gpurun_out/igemm_check.log; grep -E "FAIL|exit code" gpurun_out/igemm_check.log | head -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export VTB_GRAPH=1
B=tools/bench_conv
r() { echo "## $*"; $B "$@" | grep -E "graph|ctas" | grep -A1 "wgrad" ; }
r 256 22 22 128 128 3 1 1
r 256 11 11 256 256 3 1 1
r 256 6 6 512 512 3 1 1
r 256 22 22 128 128 1 1 0
unset VTB_GRAPH
timeout 600 python tools/layer_profile.py cspdarknet53 > gpurun_out/layers.txt 2>&1; grep -E "bn_bwd_fused|total|host" gpurun_out/layers.txt
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['gpu_launches'], d['final_loss'])"
