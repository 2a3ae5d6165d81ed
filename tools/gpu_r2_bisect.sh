#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for cfg in "VTB_TILE_PAR=1 VTB_STATS_DEFER=1" "VTB_TILE_PAR=0 VTB_STATS_DEFER=1" "VTB_TILE_PAR=1 VTB_STATS_DEFER=0" "VTB_TILE_PAR=0 VTB_STATS_DEFER=0"; do
  echo "=== $cfg"
  env $cfg timeout 200 $TR --master-port 29611 tests/dp_parity.py 2>&1 | grep -E "rank 0|timeout" | cut -c1-200 | sort | uniq -c | head -8
done
