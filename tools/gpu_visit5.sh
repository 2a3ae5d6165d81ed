#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in vovnet99_ese; do
  echo "=== $m"; timeout 600 python bench.py --model $m --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_$m.log 2>&1; tail -1 gpurun_out/bench_$m.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['gpu_launches'], d['final_loss'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})" || tail -15 gpurun_out/bench_$m.log
done
