#!/bin/bash
# 2-GPU A/B of the SyncBN exchange protocol: flags (default) vs self-certifying tagged slots
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
echo "=== dp_check tagged"; VTB_SYNC_TAGGED=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "^rank 0|FAIL|rror|timeout" | head -8 | tee gpurun_out/dp_check_tagged.log
b() { tag=$1; shift; echo "=== bench N=$N $tag"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp_$tag.err | tee gpurun_out/bench_dp${N}_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['final_loss'])"; grep -iE "error|timeout" gpurun_out/bench_dp_$tag.err | head -3; }
b tagged VTB_SYNC_TAGGED=1
b flags VTB_SYNC_TAGGED=0
