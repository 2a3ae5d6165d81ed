#!/bin/bash
# ncu --set full captures of the conv kernels in their end-of-round state: (a) tensor-bound 3x3 128->128 fprop with the
# tcgen05 counters, (b) HBM-bound 1x1 64->64 @88^2 fprop, (c) the 3x3 128->128 weight gradient.  $1 = tag
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-final}
TC="sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tmem.sum,sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,gpu__time_duration.sum"
timeout 400 ncu --set full --metrics $TC --clock-control none --import-source on -k regex:conv_igemm -c 1 -f -o gpurun_out/r02_${T}_prof_conv3x3_128 tools/bench_conv 256 22 22 128 128 3 1 1 1 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 400 ncu --set full --metrics $TC --clock-control none --import-source on -k regex:conv_igemm -c 1 -f -o gpurun_out/r02_${T}_prof_conv1x1_64 tools/bench_conv 256 88 88 64 64 1 1 0 1 >> gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 400 ncu --set full --metrics $TC --clock-control none --import-source on -k regex:wgrad_igemm -c 1 -f -o gpurun_out/r02_${T}_prof_wgrad3x3_128 tools/bench_conv 256 22 22 128 128 3 1 1 1 >> gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ls -la gpurun_out/r02_${T}_prof_*.ncu-rep
