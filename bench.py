#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native ConvNormAct/Darknet/VoVNet path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model cspdarknet53 ...]

Metric (BASELINE.json): train images/sec, CSPDarknet-53, 176 px, batch 256 per GPU, bf16, synthetic
ImageNet-shaped data, random-init weights.  One "step" = forward + loss + backward + SGD update of
backbone + classifier head (reference classifier.py:59-64, 83-95, 141-169), data-parallel with SyncBN for N > 1.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` includes the pinned-host -> device
copy of every step's images/labels and a device -> host read of the loss.  `roofline` describes the dominant
kernel (the tcgen05 implicit-GEMM convolution), `cpu_baseline` times the CPU oracle port of the same step.
`--impl reference` times the reference's CPU path (oracle port; the reference is pure Python over torch CPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

MODELS = {
    # name: (family, factory args, train GFLOP/img at `res` from BASELINE.md §2, default res, default batch/GPU)
    "cspdarknet53": ("darknet", 17.83, 176, 256),
    "darknet53": ("darknet", 27.16, 176, 256),
    "darknet19": ("darknet", 16.36, 224, 256),
    "vovnet99_ese": ("vovnet", 103.09, 224, 128),
    "darknet_yolov5l": ("yolov5", 67.27 * 3, 640, 32),   # forward 67.27 GFLOP/img (BASELINE config C5 is inference)
}
# BASELINE.json `configs`, in order: --config C1..C5 reproduces each as a driver-style line.
CONFIGS = {
    "C1": dict(model="darknet19", batch=8, res=224, precision="fp32"),     # the reference's own CPU-runnable case
    "C2": dict(model="darknet53", batch=256, res=176),
    "C3": dict(model="cspdarknet53", batch=128, res=176),                  # global 1024 at 8 GPUs, as written
    "C4": dict(model="vovnet99_ese", batch=128, res=224),
    "C5": dict(model="darknet_yolov5l", batch=32, res=640, eval=True),
}


# element-wise BatchNorm entry points reported in roofline.split: full-tensor passes (reads + writes) per call
EW_PASSES = {"vtb_bn_act": 2, "vtb_bn_bwd_fused": 5, "vtb_bn_bwd_apply": 3, "vtb_bn_bwd_reduce": 2}


def build_model(name: str):
    from vision_toolbox_b200 import backbones

    return getattr(backbones, name)()


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "gbs": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


class ProfilingLib:
    """Proxy over the ctypes library that brackets every C-ABI call with CUDA events on the launching stream."""

    def __init__(self, real):
        self._real, self.records = real, []

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if not name.startswith("vtb_") or name in ("vtb_last_error", "vtb_conv_stats_rows", "vtb_bn_bwd_rows",
                                                   "vtb_conv_wgrad_workspace_bytes", "vtb_launch_count", "vtb_pack_job_blocks", "vtb_conv_dgrad_stats_rows", "vtb_conv_dgrad_panel_w",
                                                   "vtb_conv_tiling_info", "vtb_bn_bwd_fused_rows", "vtb_sgd_job_blocks", "vtb_version",
                                                   "vtb_num_sms", "vtb_bn_sync_buffer_bytes"):
            return fn

        def wrapped(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.records.append((name, args[0] if name.startswith("vtb_conv_") else None, e0, e1, args))
            return rc

        return wrapped


def conv_flops(geom, cin_real=None) -> float:
    g = geom._obj if hasattr(geom, "_obj") else geom
    ho = (g.h + 2 * g.pad - g.k) // g.stride + 1
    wo = (g.w + 2 * g.pad - g.k) // g.stride + 1
    cin = g.cin if g.cin > 16 or cin_real is None else cin_real
    return 2.0 * g.n * ho * wo * g.cout * g.k * g.k * cin


def conv_bytes(geom, name: str) -> float:
    """Algorithmic HBM bytes of one conv launch: read the bf16 input view once, write the bf16
    output once, read the bf16 weights; wgrad reads both activations and writes the fp32 weight gradient."""
    g = geom._obj if hasattr(geom, "_obj") else geom
    ho = (g.h + 2 * g.pad - g.k) // g.stride + 1
    wo = (g.w + 2 * g.pad - g.k) // g.stride + 1
    a_in, a_out, w = g.n * g.h * g.w * g.cin * 2.0, g.n * ho * wo * g.cout * 2.0, g.cout * g.k * g.k * g.cin * 2.0
    if name.startswith("vtb_conv_wgrad"):
        return a_in + a_out + 2.0 * w
    return a_in + a_out + w


def reference_arm(args, rank: int, world: int) -> None:
    """CPU reference path: the oracle port of the training step on the host cores (bounded sample)."""
    if rank != 0:
        return
    from oracle import vt_oracle as O

    # all host cores, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for its workers)
    torch.set_num_threads(os.cpu_count() or 1)
    fam, gflop, res, _ = MODELS[args.model]
    res = args.res or res
    nb = args.cpu_batch
    torch.manual_seed(0)
    model = build_model(args.model)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    head_w = (torch.randn(1000, model.out_channels_list[-1]) * 0.01).requires_grad_(True)
    head_b = torch.zeros(1000, requires_grad=True)
    plist = list(params.values()) + [head_w, head_b]
    mom = [torch.zeros_like(p) for p in plist]
    x = torch.rand(nb, 3, res, res)
    y = torch.randint(0, 1000, (nb,))

    def step():
        if args.eval:
            with torch.no_grad():
                return float(O.features(fam, sd, x, False, "fp32")[-1].mean())
        new_stats = {}
        loss = O.classifier_loss(fam, sd, head_w, head_b, x, y, "fp32", 0.1, new_stats)
        grads = torch.autograd.grad(loss, plist)
        with torch.no_grad():
            for p, g, m in zip(plist, grads, mom):
                m.mul_(0.9).add_(g)
                p.add_(m, alpha=-0.05)
            for k, v in new_stats.items():
                sd[k] = v
        return float(loss.detach())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = nb / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "inference images/sec" if args.eval else "train images/sec", "value": val, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} {'get_feature_maps (eval)' if args.eval else 'train step'} {res}px, CPU sample "
                               f"batch {nb}", "model": args.model, "preset": args.config or None,
                   "threads": f"torch.set_num_threads(os.cpu_count()) = {cores}"},
        "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {nb} @ {res}px through oracle/vt_oracle.py (torch CPU fp32)"},
        "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(model_name: str, res: int, seconds: float = 15.0) -> dict:
    from oracle import vt_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    fam = MODELS[model_name][0]
    torch.manual_seed(0)
    model = build_model(model_name)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    head_w = (torch.randn(1000, model.out_channels_list[-1]) * 0.01).requires_grad_(True)
    head_b = torch.zeros(1000, requires_grad=True)
    nb = 8
    x, y = torch.rand(nb, 3, res, res), torch.randint(0, 1000, (nb,))

    def step():
        loss = O.classifier_loss(fam, sd, head_w, head_b, x, y, "fp32", 0.1, {})
        torch.autograd.grad(loss, params + [head_w, head_b])

    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step(); n += 1
        if time.perf_counter() - t0 > seconds or n >= 20:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": nb / dt, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} fwd+bwd steps of batch {nb} @ {res}px, oracle/vt_oracle.py on torch CPU fp32"}


def eval_arm(args, rank: int, world: int, local_rank: int) -> None:
    """Inference arm (BASELINE config C5): model.get_feature_maps(x) under no_grad, eval mode (BatchNorm folded into the
    conv epilogue: one launch per ConvNormAct).  N > 1: independent replicas, one per GPU (no exchange step exists)."""
    import torch.distributed as dist

    from vision_toolbox_b200 import _lib

    fam, gflop_img, dres, dbatch = MODELS[args.model]
    res, nb = args.res or dres, args.batch or dbatch
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(args.warmup, 3), args.steps
    torch.manual_seed(0)
    model = build_model(args.model).to(dev).eval()
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_x = [torch.rand(nb, 3, res, res, generator=g).pin_memory() for _ in range(2)]
    dev_x = [h.to(dev) for h in host_x]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(x):
        with torch.no_grad():
            return model.get_feature_maps(x)

    for i in range(W):
        step(dev_x[i % 2])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        outs = step(dev_x[i % 2])
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    value = nb * world * K / (float(ms) / 1e3)
    # end to end: pinned host -> device copy of every batch, a scalar of the last map read back
    buf = torch.empty_like(dev_x[0])
    barrier()
    e0.record()
    for i in range(K):
        buf.copy_(host_x[i % 2], non_blocking=True)
        last = float(step(buf)[-1].float().mean())
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    roofline = None
    if rank == 0:
        prof = ProfilingLib(_lib.lib())
        runners = list(model.__dict__.get("_vtb_plans", {}).values())
        for r in runners:
            r.L = prof
        torch.cuda._sleep(30_000_000)
        step(dev_x[0])
        torch.cuda.synchronize()
        for r in runners:
            r.L = _lib.lib()
        pk = peaks()
        t = fl = n = 0
        for name, geom, a, b, _ in prof.records:
            if name == "vtb_conv_fprop":
                t += a.elapsed_time(b); fl += conv_flops(geom); n += 1
        ach = fl / (t / 1e3) / 1e12 if t else 0.0
        roofline = {"bound": "tensor", "kernel": "conv_igemm_kernel (eval fprop, fused BN+ReLU epilogue)", "achieved": ach,
                    "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": None,
                    "peak_source": pk["source"] + ", sustained cuBLAS bf16", "launches": n, "avg_launch_ms": t / max(n, 1)}
        line = {"metric": "inference images/sec", "value": value, "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": float(ms) / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{args.model} get_feature_maps (eval, no_grad), {res}px, batch {nb}/GPU, bf16",
                           "model": args.model, "global_batch": nb * world, "resolution": res,
                           "parallelism": f"replicas x{world}", "preset": args.config or None,
                           "l2_policy": "activations >> L2"},
                "clocks": clocks,
                "e2e": {"value": nb * world * K / (float(ms2) / 1e3), "unit": "img/s",
                        "h2d_bytes_per_step": host_x[0].numel() * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
                "model_tflops": value * gflop_img / 3 / 1e3, "last": last}
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="cspdarknet53", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sync-bn", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step from a CUDA graph (single GPU)")
    ap.add_argument("--config", default="", choices=[""] + list(CONFIGS), help="BASELINE.json config preset (C1..C5)")
    ap.add_argument("--eval", action="store_true", help="inference: get_feature_maps under no_grad (config C5)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    if args.config:
        preset = CONFIGS[args.config]
        args.model = preset["model"]
        args.batch = args.batch or preset["batch"]
        args.res = args.res or preset["res"]
        args.eval = args.eval or preset.get("eval", False)
        args.precision = preset.get("precision", args.precision)
        if args.impl == "reference":
            args.cpu_batch = min(args.cpu_batch, args.batch)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.eval:
        eval_arm(args, rank, world, local_rank)
        return

    import torch.distributed as dist

    from vision_toolbox_b200 import _lib, parallel

    fam, gflop_img, dres, dbatch = MODELS[args.model]
    res, nb = args.res or dres, args.batch or dbatch
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the SyncBN kernels leave VTB_SM_RESERVE (16) SMs free for the gradient all-reduce: keep NCCL inside them
        from vision_toolbox_b200.parallel import nccl_pg_options

        opts = nccl_pg_options()
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    W = max(args.warmup, 3)
    K = args.steps

    torch.manual_seed(0)
    if args.precision == "fp32":
        import vision_toolbox_b200

        vision_toolbox_b200.set_precision("fp32")   # config C1: the fp32 parity kernels (CUDA cores), not the fast path
    model = build_model(args.model).to(dev).train()
    head = torch.nn.Linear(model.out_channels_list[-1], 1000).to(dev)
    trainer = parallel.Trainer(model, head, lr=0.05, momentum=0.9, weight_decay=2e-5, label_smoothing=0.1,
                               sync_bn=not args.no_sync_bn, process_group=dist.group.WORLD if world > 1 else None)

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_host = 2
    host_x = [torch.rand(nb, 3, res, res, generator=g).pin_memory() for _ in range(n_host)]
    host_y = [torch.randint(0, 1000, (nb,), generator=g).pin_memory() for _ in range(n_host)]
    dev_x = [h.to(dev) for h in host_x]
    dev_y = [h.to(dev) for h in host_y]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    use_graph = not args.no_graph and (world == 1 or os.environ.get("VTB_DP_GRAPH", "1") == "1")
    if use_graph:
        try:
            trainer.enable_cuda_graph(dev_x[0], dev_y[0], warmup=2)
        except Exception as e:  # noqa: BLE001 - any capture problem -> eager replay of the same kernels
            if rank == 0:
                print(f"[bench] CUDA graph disabled: {type(e).__name__}: {e}", file=sys.stderr)
            trainer._graph = None
            use_graph = False
        if world > 1:   # every rank must take the same path
            flag = torch.tensor([1 if use_graph else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0:
                trainer._graph = None
                use_graph = False
    for i in range(W):
        trainer.step(dev_x[i % n_host], dev_y[i % n_host])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = trainer.step(dev_x[i % n_host], dev_y[i % n_host])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    launches = _lib.launch_count() - launches0
    if use_graph:
        launches += K * trainer.graph_launches   # replayed launches do not pass through the host-side counter
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    clocks = sampler.stop() if rank == 0 else None
    value = nb * world * K / (ms_total / 1e3)
    final_loss = float(loss)

    # ---------------- end to end: pinned host -> device every step, loss read back every step ----------------
    copy_stream = torch.cuda.Stream(dev)
    bufs_x = [torch.empty_like(dev_x[0]) for _ in range(2)]
    bufs_y = [torch.empty_like(dev_y[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            bufs_x[s].copy_(host_x[i % n_host], non_blocking=True)
            bufs_y[s].copy_(host_y[i % n_host], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(2):
            consumed[s].record()
        prefetch(0)
        out = 0.0
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            l = trainer.step(bufs_x[s], bufs_y[s])
            consumed[s].record()
            out = float(l)  # device -> host read of the step's result
        return out

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(K)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = nb * world * K / (float(ms2) / 1e3)
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8

    # ---------------- roofline of the dominant kernel (instrumented extra step, rank 0) ----------------
    roofline, breakdown = None, None
    # every rank runs the extra step (it contains the SyncBN / gradient collectives); rank 0 instruments it
    prof = ProfilingLib(_lib.lib())
    runners = list(model.__dict__.get("_vtb_plans", {}).values())
    sides = [getattr(r, "_side", None) for r in runners]
    for r in runners:
        r._side = None          # instrumented step: one stream, so that every event pair brackets exactly one kernel family
    for r in runners:
        r.L = prof       # EVERY rank runs the instrumented step (same host pacing on all ranks); rank 0 reports it
    barrier()
    torch.cuda._sleep(30_000_000)  # let the host run ahead so event intervals contain no launch gaps
    real_lib_fn = _lib.lib
    _lib.lib = lambda: prof      # the head / loss / optimizer entry points look the library up at call time
    try:
        trainer._step_eager(dev_x[0], dev_y[0])   # eager on purpose: per-call events cannot be recorded inside a graph replay
    finally:
        _lib.lib = real_lib_fn
    barrier()
    for r, sd_ in zip(runners, sides):
        r.L = _lib.lib()
        r._side = sd_
    if rank == 0:
        agg = {}
        pk = peaks()
        # every conv_igemm launch is also classified by the roofline that bounds ITS shape (algorithmic flops / bytes
        # against the measured peaks): the 3x3 layers with >= 128 channels are tensor-bound, stems / 1x1 / narrow
        # layers are HBM-bound; "roofline" below is the kernel as a whole, "roofline.split" the two classes apart
        split = {"tensor": [0.0, 0.0, 0.0, 0], "hbm": [0.0, 0.0, 0.0, 0]}   # ms, flops, bytes, launches
        ew = {}   # element-wise BatchNorm passes: entry point -> [ms, algorithmic bytes, launches]
        for name, geom, a, b, cargs in prof.records:
            t = a.elapsed_time(b)
            if name in EW_PASSES:
                # (pixels, channels) sit at the same argument positions in all of them; passes = algorithmic full-tensor
                # reads + writes of the call
                pix, ch = cargs[4], cargs[5]
                if name == "vtb_bn_act":
                    pix, ch = cargs[2], cargs[3]
                    passes = 3 if cargs[7] else 2
                else:
                    passes = EW_PASSES[name]
                e = ew.setdefault(name, [0.0, 0.0, 0])
                e[0] += t; e[1] += float(pix) * ch * 2.0 * passes; e[2] += 1
            is_conv = geom is not None and name in ("vtb_conv_fprop", "vtb_conv_fprop_bn", "vtb_conv_dgrad", "vtb_conv_dgrad_bn", "vtb_conv_wgrad", "vtb_conv_wgrad_pair")
            fl = conv_flops(geom) if is_conv else 0.0
            d = agg.setdefault(name, [0.0, 0.0, 0])
            d[0] += t; d[1] += fl; d[2] += 1
            if is_conv and not name.startswith("vtb_conv_wgrad"):
                by = conv_bytes(geom, name)
                cls = "tensor" if fl / (pk["tflops"] * 1e12) >= by / (pk["gbs"] * 1e9) else "hbm"
                c = split[cls]
                c[0] += t; c[1] += fl; c[2] += by; c[3] += 1
        conv_calls = [agg.get(k, [0, 0, 0]) for k in ("vtb_conv_fprop", "vtb_conv_fprop_bn", "vtb_conv_dgrad", "vtb_conv_dgrad_bn")]
        igemm_ms = sum(v[0] for v in conv_calls)
        igemm_fl = sum(v[1] for v in conv_calls)
        igemm_n = sum(v[2] for v in conv_calls)
        achieved = igemm_fl / (igemm_ms / 1e3) / 1e12 if igemm_ms > 0 else 0.0
        traffic = None   # dram__bytes_read+write per launch from the committed ncu pass (tools/launches_summary.py)
        tp = ROOT / "profiles" / "ncu_conv_traffic.json"
        if tp.exists() and args.model == "cspdarknet53" and nb == 256 and res == 176:
            traffic = json.loads(tp.read_text()).get("avg_dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "conv_igemm_kernel (fprop + dgrad launches)", "achieved": achieved,
                    "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic,
                    "peak_source": pk["source"] + ", sustained cuBLAS bf16", "launches": igemm_n,
                    "avg_launch_ms": igemm_ms / max(igemm_n, 1)}
        tb, hb = split["tensor"], split["hbm"]
        roofline["split"] = {
            "tensor_bound": {"launches": tb[3], "ms": round(tb[0], 3), "achieved": tb[1] / max(tb[0], 1e-9) / 1e9,
                             "peak": pk["tflops"], "unit": "TFLOP/s", "frac": tb[1] / max(tb[0], 1e-9) / 1e9 / pk["tflops"]},
            "hbm_bound": {"launches": hb[3], "ms": round(hb[0], 3), "achieved": hb[2] / max(hb[0], 1e-9) / 1e6,
                          "peak": pk["gbs"], "unit": "GB/s", "frac": hb[2] / max(hb[0], 1e-9) / 1e6 / pk["gbs"]}}
        for nm, (ms_, by_, n_) in ew.items():
            roofline["split"][nm] = {"launches": n_, "ms": round(ms_, 3), "achieved": by_ / max(ms_, 1e-9) / 1e6,
                                     "peak": pk["gbs"], "unit": "GB/s", "frac": by_ / max(ms_, 1e-9) / 1e6 / pk["gbs"]}
        tot = sum(v[0] for v in agg.values())
        breakdown = {k: {"ms": round(v[0], 3), "share": round(v[0] / tot, 3), "calls": v[2],
                         **({"tflops": round(v[1] / (v[0] / 1e3) / 1e12, 1)} if v[1] else {})}
                     for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.model, res)

    if rank == 0:
        pk = peaks()
        line = {
            "metric": "train images/sec", "value": value, "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} train step (fwd+loss+bwd+SGD), {res}px, batch {nb}/GPU, {args.precision}, "
                                   f"{'SyncBN+DDP' if world > 1 else 'single GPU'}",
                       "model": args.model, "global_batch": nb * world, "resolution": res,
                       "parallelism": f"dp{world}", "l2_policy": "inputs+activations >> L2 (multi-GB working set per step)",
                       "cuda_graph": bool(use_graph)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "model_tflops": value * gflop_img / 1e3,
            "model_frac_of_tensor_peak": value * gflop_img / 1e3 / pk["tflops"],
            "final_loss": final_loss,
            "kernel_breakdown": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        if use_graph:
            # a captured graph holds NCCL work; tearing the process group down under it can block at interpreter exit,
            # so leave without running destructors (every rank is past the barrier, rank 0 has printed its line)
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
