#!/bin/bash
# development visit: harness + parity + per-layer profile, with PDL on and off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_check_igemm.sh > /dev/null 2>&1; grep -E "FAIL|exit code|error|mismatch|timeout" gpurun_out/igemm_check.log | sort | uniq -c | sort -rn | head -20; grep -c " ok " gpurun_out/igemm_check.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/layer_profile.py ${1:-cspdarknet53} > gpurun_out/layers.txt 2>&1; head -${2:-40} gpurun_out/layers.txt
echo "=== bench PDL on"; timeout 600 python bench.py --no-cpu-baseline --steps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])"
echo "=== bench PDL off"; VTB_PDL=0 timeout 600 python bench.py --no-cpu-baseline --steps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])"
