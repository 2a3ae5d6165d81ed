#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), launch list under ncu. Outputs -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python tools/one_step.py cspdarknet53 256 176 2 > gpurun_out/ncu_run.log 2>&1; tail -3 gpurun_out/ncu_run.log
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json
echo "=== ncu full: 128->128 3x3 fprop/dgrad (tools/bench_conv)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 2 -f -o gpurun_out/prof_conv3x3_128 tools/bench_conv 256 22 22 128 128 3 1 1 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_igemm -c 1 -f -o gpurun_out/prof_wgrad3x3_128 tools/bench_conv 256 22 22 128 128 3 1 1 >> gpurun_out/ncu_full.log 2>&1
timeout 300 python tools/layer_profile.py cspdarknet53 > gpurun_out/layers.txt 2>&1; head -12 gpurun_out/layers.txt
