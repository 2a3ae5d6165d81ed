"""Summarise an .ncu-rep (one block per profiled launch) into the handful of numbers the roofline uses.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.txt   (runs on the CPU box: ncu -i)"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    # tcgen05 work (round 2: the counters that DO register UTCHMMA; the hmma / pipe_tensor ones below are legacy-HMMA)
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.per_cycle_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.peak_sustained",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_reads.sum", "sm__mem_tensor_writes.sum", "sm__inst_executed_pipe_tmem.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        print("kernel:", d.get("Kernel Name", ("?",))[0][:150])
        for k in KEYS:
            hit = [h for h in hdr if h == k or h.endswith("." + k)]
            for h in hit[:1]:
                print(f"  {k:80s} {d[h][0]:>16s} {d[h][1]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
