#!/bin/bash
# end-of-round validation on one GPU: tests, smoke, headline bench (+ CPU baseline), reference arm, every BASELINE config
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-final}
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_$T.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke_$T.log
echo "=== bench (default flags)"; timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_$T.json | cut -c1-300; tail -2 gpurun_out/bench.err
echo "=== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/r02_bench_ref_$T.json | cut -c1-300; tail -2 gpurun_out/bench_ref.err
echo "=== configs"; : > gpurun_out/r02_bench_configs_$T.jsonl
for c in C2 C3 C4 C5; do timeout 600 python bench.py --config $c --no-cpu-baseline 2>> gpurun_out/bench.err | tee -a gpurun_out/r02_bench_configs_$T.jsonl | cut -c1-160; done
echo done
