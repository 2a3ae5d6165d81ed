#!/bin/bash
# Round-2 visit 1: full GPU parity suite (incl. the benchmark-regime cases), baseline bench on this box, the torch-eager
# cuDNN comparator for every BASELINE config, compute-sanitizer on the conv kernels.  Outputs -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 | tee gpurun_out/r02_pytest_gpu_t1.log | tail -15
echo "=== bench (headline)"; timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/r02_bench_t1.json | cut -c1-600; tail -3 gpurun_out/bench.err
echo "=== torch eager cuDNN comparator"
for cfg in "cspdarknet53 256 176" "cspdarknet53 128 176" "darknet53 256 176" "vovnet99_ese 128 224"; do
  set -- $cfg
  timeout 300 python tools/bench_torch_eager.py --model $1 --batch $2 --res $3 --steps 10 --warmup 5 2>&1 | tail -1 | tee -a gpurun_out/r02_torch_eager.jsonl | cut -c1-300
done
timeout 300 python tools/bench_torch_eager.py --model darknet_yolov5l --batch 32 --res 640 --eval --steps 10 --warmup 5 2>&1 | tail -1 | tee -a gpurun_out/r02_torch_eager.jsonl | cut -c1-300
echo "=== ours, other configs"
for c in C2 C3 C4 C5; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench.err | tee -a gpurun_out/r02_bench_configs_t1.jsonl | cut -c1-400; done
echo "=== compute-sanitizer memcheck (tools/test_igemm cases 0 2 10)"
for i in 0 2 10; do timeout 200 compute-sanitizer --tool memcheck tools/test_igemm $i 2>&1 | tail -6; done | tee gpurun_out/r02_sanitizer_memcheck.log | tail -20
echo "=== compute-sanitizer racecheck (case 2)"
timeout 240 compute-sanitizer --tool racecheck tools/test_igemm 2 2>&1 | tail -8 | tee gpurun_out/r02_sanitizer_racecheck.log
echo done
