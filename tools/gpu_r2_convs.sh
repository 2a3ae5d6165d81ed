#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { VTB_GRAPH=1 timeout 120 tools/bench_conv $1 10 2>&1 | grep -E "graph replay" | grep -v "dgradbn\|dgr+acc" | sed 's/(host-free, back to back)//g; s/graph replay://g; s/per launch//g' | tr '\n' ' '; }
{
for g in "128 7 7 2144 1024 1 1 0" "128 7 7 1888 1024 1 1 0" "128 14 14 512 192 3 1 1"; do
  for bm in 0 128; do VTB_BLOCK_M=$bm run "$g"; echo " <- block_m=$bm $g"; done
done
} 2>&1 | tee gpurun_out/r02_convs_vovnet_tiles2.txt
echo "=== bench C4 (vovnet99)"; timeout 300 python bench.py --config C4 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_C4_t13.json | cut -c1-200; tail -2 gpurun_out/bench.err
echo "=== bench default"; timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t13.json | cut -c1-200; tail -2 gpurun_out/bench.err
