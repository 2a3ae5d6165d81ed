#!/bin/bash
# multi-GPU visit (gpurun --gpus N): DP parity + bench at N GPUs with and without the CUDA graph
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
echo "skip dp_check"
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}_$2.log 2>&1; grep '^{' gpurun_out/bench_${N}_$2.log | tail -1 > gpurun_out/bench_${N}_$2.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_${N}_$2.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('cuda_graph'), d['final_loss'])" || tail -30 gpurun_out/bench_${N}_$2.log; }
echo "=== bench $N GPU graph"; run 29513 graph

