#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== bench ours ($N GPUs)"; timeout 300 $TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp.err | tee gpurun_out/r02_bench_${N}gpu.json | cut -c1-200; tail -2 gpurun_out/bench_dp.err
echo "=== reference arm under torchrun ($N ranks; rank 0 works)"; timeout 300 $TR --master-port 29603 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2> gpurun_out/bench_ref_dp.err | tee gpurun_out/r02_bench_ref_${N}gpu.json | cut -c1-260; tail -2 gpurun_out/bench_ref_dp.err
echo done
