#!/bin/bash
# GPU-box visit: full GPU parity suite (bf16 + fp32 mode, no -x: every failure is reported), bench (both arms),
# launch list under ncu WITH dram bytes (roofline.traffic), per-layer profile.  Outputs -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -120 > gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python tools/one_step.py cspdarknet53 256 176 2 > gpurun_out/ncu_run.log 2>&1; tail -3 gpurun_out/ncu_run.log
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json
timeout 300 python tools/layer_profile.py cspdarknet53 > gpurun_out/layers.txt 2>&1; head -12 gpurun_out/layers.txt
