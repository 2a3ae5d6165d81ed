#!/bin/bash
# ncu launch list (durations + DRAM bytes) of ONE training step.  $2 = cache control (all: flush before every kernel, ncu's
# default; none: keep the caches as the previous kernel left them)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-t13}
CC=${2:-all}
timeout 900 ncu --cache-control $CC --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_ncu_launches_step_$T.csv python tools/one_step.py cspdarknet53 256 176 2 > gpurun_out/ncu_run.log 2>&1; tail -2 gpurun_out/ncu_run.log
python tools/launches_summary.py gpurun_out/r02_ncu_launches_step_$T.csv > gpurun_out/r02_ncu_launches_step_${T}_summary.txt 2>&1; head -16 gpurun_out/r02_ncu_launches_step_${T}_summary.txt
echo done
