#!/bin/bash
# N-GPU sanity: data-parallel parity + bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
echo "=== dp_check N=$N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep -E "^rank 0|FAIL|rror" | head -8 | tee gpurun_out/dp_check_$N.log
echo "=== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_dp$N.err | tee gpurun_out/bench_dp${N}_v6.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d['config']['cuda_graph'], d['e2e']['value'])"; grep -iE "error|timeout|Traceback" gpurun_out/bench_dp$N.err | head -5
