#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== bench graph"; timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_graph.log 2>&1; tail -1 gpurun_out/bench_graph.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['gpu_launches'], d['final_loss'])" || tail -20 gpurun_out/bench_graph.log
echo "=== bench eager"; timeout 600 python bench.py --no-cpu-baseline --steps 20 --no-graph 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['gpu_launches'], d['final_loss'])"
