#!/bin/bash
# N-GPU visit: DP parity + bench (fused peer-memory SyncBN)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
echo "=== dp_check"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "SyncBN via|Error|error" | head -12
echo "=== bench $N GPU"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; grep '^{' gpurun_out/bench_${N}gpu.log | tail -1 > gpurun_out/bench_${N}gpu.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_${N}gpu.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})" || tail -30 gpurun_out/bench_${N}gpu.log
