#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu_latest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== layers vovnet99_ese"; timeout 300 python tools/layer_profile.py vovnet99_ese 128 224 > gpurun_out/r02_layers_vovnet99.txt 2>&1; head -30 gpurun_out/r02_layers_vovnet99.txt | cut -c1-150
echo "=== layers darknet53"; timeout 300 python tools/layer_profile.py darknet53 256 176 > gpurun_out/r02_layers_darknet53.txt 2>&1; head -12 gpurun_out/r02_layers_darknet53.txt | cut -c1-150
echo done
