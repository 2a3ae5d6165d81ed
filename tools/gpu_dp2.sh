#!/bin/bash
# 2-GPU visit: data-parallel parity (bf16 p2p/nccl SyncBN, fp32 mode) + 2-GPU bench with / without side-stream wgrad
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
echo "=== dp_check"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "^rank 0|Error|error|FAIL" | head -20 | tee gpurun_out/dp_check.log
echo "=== dp_check VTB_WGRAD_STREAM=1"; VTB_WGRAD_STREAM=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py 2>&1 | grep -E "^rank 0|Error|error|FAIL" | head -20 | tee gpurun_out/dp_check_ws.log
echo "=== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp.err | tee gpurun_out/bench_dp$N.json | cut -c1-260; tail -3 gpurun_out/bench_dp.err
echo "=== bench N=$N VTB_WGRAD_STREAM=1"; VTB_WGRAD_STREAM=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp_ws.err | tee gpurun_out/bench_dp${N}_ws.json | cut -c1-260; tail -3 gpurun_out/bench_dp_ws.err
