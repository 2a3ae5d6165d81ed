#!/bin/bash
# multi-rank A/B: weight-gradient side stream on/off, SM reserve on/off (N ranks, default 2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for cfg in "VTB_WGRAD_STREAM_MULTI=0 VTB_SM_RESERVE=16" "VTB_WGRAD_STREAM_MULTI=1 VTB_SM_RESERVE=16" "VTB_WGRAD_STREAM_MULTI=1 VTB_SM_RESERVE=0" "VTB_WGRAD_STREAM_MULTI=0 VTB_SM_RESERVE=0"; do
  i=$((i+1))
  echo "=== $cfg"
  env $cfg timeout 150 $TR --master-port $((29620+i)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/side_$i.err | tee gpurun_out/r02_side_${N}gpu_$i.json | cut -c1-200
  tail -2 gpurun_out/side_$i.err | cut -c1-300
done
echo done
