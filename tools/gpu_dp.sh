#!/bin/bash
# multi-GPU visit (gpurun --gpus N): DP parity (peer-memory SyncBN and NCCL SyncBN) + bench at 1 and N GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m 2>&1 | head -12
echo "=== dp_check p2p"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -v -E "^\s*$|OMP_NUM|\*\*\*\*" | tail -12
echo "=== dp_check nccl"; VTB_SYNCBN=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py 2>&1 | grep -v -E "^\s*$|OMP_NUM|\*\*\*\*" | tail -6
echo "=== bench 1 GPU"; timeout 600 python bench.py --no-cpu-baseline --steps 20 2>&1 | tail -1 | tee gpurun_out/bench_1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "=== bench $N GPU p2p"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep '^{' | tail -1 | tee gpurun_out/bench_$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "=== bench $N GPU nccl-syncbn"; VTB_SYNCBN=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
