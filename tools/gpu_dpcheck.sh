#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-4}
echo "=== dp_check p2p"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "SyncBN via" | head -3
echo "=== dp_check nccl"; VTB_SYNCBN=nccl timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py 2>&1 | grep -E "SyncBN via" | head -3
