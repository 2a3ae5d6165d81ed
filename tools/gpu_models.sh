#!/bin/bash
# every benchmark family once: a few training steps (finite loss), throughput line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in darknet53 vovnet99_ese darknet19; do
  echo "=== $m"; timeout 600 python bench.py --model $m --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_$m.log 2>&1; tail -1 gpurun_out/bench_$m.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['gpu_launches'], d['final_loss'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})" || tail -15 gpurun_out/bench_$m.log
done
echo "=== yolov5l eval 32x640"; timeout 300 python - <<'PY'
import torch, time
from vision_toolbox_b200 import backbones
m = backbones.darknet_yolov5l().cuda().eval()
x = torch.rand(32, 3, 640, 640, device="cuda")
with torch.no_grad():
    for _ in range(3): f = m.get_feature_maps(x)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f = m.get_feature_maps(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print([tuple(t.shape) for t in f], f"{ms:.2f} ms/iter  {32/ms*1e3:.0f} img/s  {32*67.27/ms:.0f} TFLOP/s", all(torch.isfinite(t.float()).all().item() for t in f))
PY
