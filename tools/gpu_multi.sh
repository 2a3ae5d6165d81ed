#!/bin/bash
# N-rank parity against the oracle + bench line + the reference arm under torchrun (rank 0 only, all host cores)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}   # ranks = GPUs of the box: gpurun --gpus N -- bash tools/gpu_multi.sh N
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== dp_parity ($N ranks)"; timeout 300 $TR --master-port 29601 tests/dp_parity.py 2>&1 | grep -E "rank 0|FAIL|Error|error" | tee gpurun_out/r02_dp_parity_${N}gpu.log | cut -c1-250
echo "=== bench ours ($N GPUs)"; timeout 300 $TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp.err | tee gpurun_out/r02_bench_${N}gpu.json | cut -c1-200; tail -2 gpurun_out/bench_dp.err
echo "=== pytest multi-GPU test"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
echo done
