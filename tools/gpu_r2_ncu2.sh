#!/bin/bash
# ncu --set full capture of ONE fused-normalise fprop launch (1x1 128->128 @22^2, batch 256)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VTB_FUSED_NORM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 3 -c 1 -f -o gpurun_out/r02_prof_fusednorm_1x1_128 python tools/one_unit.py 128 128 1 1 22 256 5 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*fusednorm*.ncu-rep
