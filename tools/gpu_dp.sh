#!/bin/bash
# multi-GPU visit (gpurun --gpus N): [DP parity] + bench at N GPUs (peer-memory SyncBN vs NCCL SyncBN)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
if [ "$2" == "check" ]; then
echo "=== dp_check p2p"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep "SyncBN via"
fi
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$N_$2.log 2>&1; grep '^{' gpurun_out/bench_$N_$2.log | tail -1 > gpurun_out/bench_$N_$2.json; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_$N_$2.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})" || tail -30 gpurun_out/bench_$N_$2.log; }
echo "=== bench $N GPU p2p"; run 29513 p2p
echo "=== bench $N GPU nccl-syncbn"; VTB_SYNCBN=nccl run 29514 nccl
