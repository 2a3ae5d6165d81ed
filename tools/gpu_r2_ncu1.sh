#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== ncu metric names that could see tcgen05 / TMEM work"
ncu --query-metrics --chip gb100 2>/dev/null | grep -i -E "tensor|utc|tmem|tcgen|umma" | awk '{print $1}' | sort -u > gpurun_out/r02_ncu_tensor_metric_names.txt || true
wc -l gpurun_out/r02_ncu_tensor_metric_names.txt; head -80 gpurun_out/r02_ncu_tensor_metric_names.txt
ncu --list-chips 2>/dev/null | head -5
echo "=== launch list of one step"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_ncu_launches_step.csv python tools/one_step.py cspdarknet53 256 176 2 > gpurun_out/ncu_run.log 2>&1; tail -2 gpurun_out/ncu_run.log
python tools/launches_summary.py gpurun_out/r02_ncu_launches_step.csv > gpurun_out/r02_ncu_launches_step_summary.txt 2>&1; head -16 gpurun_out/r02_ncu_launches_step_summary.txt
echo done
