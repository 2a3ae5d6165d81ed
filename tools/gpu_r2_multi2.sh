#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== dp_parity ($N ranks)"; timeout 300 $TR --master-port 29601 tests/dp_parity.py 2>&1 | grep -E "rank 0|FAIL|Error|error" | tee gpurun_out/r02_dp_parity_${N}gpu.log | cut -c1-250
echo "=== bench ours ($N GPUs), optimizer overlap"; timeout 300 $TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp.err | tee gpurun_out/r02_bench_${N}gpu.json | cut -c1-200; tail -3 gpurun_out/bench_dp.err
echo "=== bench ours ($N GPUs), VTB_SGD_OVERLAP=0"; VTB_SGD_OVERLAP=0 timeout 300 $TR --master-port 29603 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp2.err | tee gpurun_out/r02_bench_${N}gpu_nooverlap.json | cut -c1-200; tail -3 gpurun_out/bench_dp2.err
echo done
