#!/bin/bash
# 2-GPU: where does the data-parallel loss come from?  NCCL channel count / SyncBN off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
b() { tag=$1; shift; echo "=== bench N=$N $tag"; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 20 --warmup 5 $EXTRA 2> gpurun_out/bench_dp_$tag.err | tee gpurun_out/bench_dp${N}_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; grep -iE "error|timeout" gpurun_out/bench_dp_$tag.err | head -3; }
echo "=== dp_check"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "^rank 0|FAIL" | head -8 | tee gpurun_out/dp_check.log
b default A=1
b ch2 NCCL_MAX_NCHANNELS=2
b ch4 NCCL_MAX_NCHANNELS=4
EXTRA=--no-sync-bn b nosyncbn A=1
