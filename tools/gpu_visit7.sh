#!/bin/bash
# final visit of the round: everything gpu_visit6 does + one ncu --set full capture of three consecutive in-step conv launches
# (1x1 / 3x3 fprop+BN of a CSP block at 22x22) + the other model families
cd "$(dirname "$0")/.."
bash tools/gpu_visit6.sh
echo "=== ncu full (in-step conv launches)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 157 --launch-count 3 -f -o gpurun_out/prof_instep_conv python tools/one_step.py cspdarknet53 256 176 2 > gpurun_out/ncu_full_instep.log 2>&1; tail -2 gpurun_out/ncu_full_instep.log
timeout 120 ncu -i gpurun_out/prof_instep_conv.ncu-rep --page details --csv > gpurun_out/prof_instep_conv_details.csv 2>/dev/null
bash tools/gpu_models.sh 2>&1 | tee gpurun_out/models_v6.txt | grep -vE "^\s*$" | tail -8
