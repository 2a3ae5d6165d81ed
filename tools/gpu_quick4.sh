#!/bin/bash
# A/B on one box: env switch given as arguments, e.g.  tools/gpu_quick4.sh VTB_BWD_COOP=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -15
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench_$tag.err | tee gpurun_out/bench_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value']), round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"; tail -2 gpurun_out/bench_$tag.err; }
run new A=1
run alt "$@"
run new2 A=1
