#!/bin/bash
# quick GPU check: parity suite + smoke + bench (pair on / off)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench_quick.json; tail -5 gpurun_out/bench.err
echo "=== bench VTB_PAIR=0"; VTB_PAIR=0 timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench_nopair.err | tee gpurun_out/bench_nopair.json; tail -5 gpurun_out/bench_nopair.err
