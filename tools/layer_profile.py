"""Per-call device timing of one training step (CUDA events around every C-ABI call), grouped by conv shape.
    python tools/layer_profile.py [model] [batch] [res]
Prints, per (entry point, geometry): calls, ms, TFLOP/s and algorithmic GB/s; plus host enqueue time per step."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from bench import ProfilingLib, conv_flops
from vision_toolbox_b200 import _lib, backbones, parallel

name = sys.argv[1] if len(sys.argv) > 1 else "cspdarknet53"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
res = int(sys.argv[3]) if len(sys.argv) > 3 else 176
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = getattr(backbones, name)().to(dev).train()
head = torch.nn.Linear(model.out_channels_list[-1], 1000).to(dev)
tr = parallel.Trainer(model, head)
x = torch.rand(nb, 3, res, res, device=dev)
y = torch.randint(0, 1000, (nb,), device=dev)
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()

t0 = time.perf_counter()
tr.step(x, y)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {t_host*1e3:.2f} ms, step wall {t_all*1e3:.2f} ms")

prof = ProfilingLib(_lib.lib())
runners = list(model.__dict__.get("_vtb_plans", {}).values())
for r in runners:
    r.L = prof
    r._side = None   # single stream: every event pair brackets exactly one call
torch.cuda.synchronize()
torch.cuda._sleep(60_000_000)
tr.step(x, y)
torch.cuda.synchronize()
agg = {}
for fn, geom, a, b, args in prof.records:
    t = a.elapsed_time(b)
    key = fn
    fl = by = 0.0
    if geom is not None and fn in ("vtb_conv_fprop", "vtb_conv_fprop_bn", "vtb_conv_dgrad", "vtb_conv_dgrad_bn", "vtb_conv_wgrad", "vtb_conv_wgrad_pair"):
        g = geom._obj
        ho = (g.h + 2 * g.pad - g.k) // g.stride + 1
        key = f"{fn[9:]:8s} {g.cin:4d}->{g.cout:4d} k{g.k}s{g.stride} {g.h:3d}->{ho:3d}"
        fl = conv_flops(geom)
        by = 2.0 * g.n * (g.h * g.w * g.cin + ho * ho * g.cout) + 2.0 * g.k * g.k * g.cin * g.cout
    elif fn in ("vtb_bn_act", "vtb_bn_bwd_reduce", "vtb_bn_bwd_apply", "vtb_bn_bwd_fused"):
        # args: (y/dout, ld, pixels, c, ...) resp. (dout, lddo, y, ldy, pixels, c, ...)
        pix, c = (args[2], args[3]) if fn == "vtb_bn_act" else (args[4], args[5])
        key = f"{fn[4:]:14s} c{c:4d} pix{pix:8d}"
        passes = {"vtb_bn_act": 2 + (1 if args[7] else 0), "vtb_bn_bwd_reduce": 2, "vtb_bn_bwd_apply": 3,
                  "vtb_bn_bwd_fused": 5}[fn]
        by = 2.0 * pix * c * passes
    d = agg.setdefault(key, [0.0, 0.0, 0.0, 0])
    d[0] += t; d[1] += fl; d[2] += by; d[3] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    extra = f" {v[1]/v[0]/1e9:7.1f} TF/s {v[2]/v[0]/1e6:7.1f} GB/s" if v[1] else (f" {v[2]/v[0]/1e6:19.1f} GB/s" if v[2] else "")
    print(f"{v[0]:8.3f} ms {100*v[0]/tot:5.1f}% n={v[3]:3d} {k}{extra}")
