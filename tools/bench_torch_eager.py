#!/usr/bin/env python
"""The bar to beat on the same box: the reference's module composition under torch-eager CUDA (cuDNN), NOT a product path.

BASELINE.md section 3 / reference configs/base.yaml:16-23: autocast(bf16), channels_last, cudnn.benchmark, SGD(momentum,
weight decay on conv/linear weights), label-smoothed CE; DDP + SyncBatchNorm when launched under torchrun.  The module tree
is the package's own (identical to the reference's, state_dict-compatible); this script re-binds every native-dispatching
``forward`` to the plain torch composition the modules keep for CPU tensors, so that CUDA tensors run ATen/cuDNN kernels.

    python tools/bench_torch_eager.py --model cspdarknet53 --batch 256 --res 176
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_torch_eager.py --model ...
    python tools/bench_torch_eager.py --model darknet_yolov5l --batch 32 --res 640 --eval

Prints one JSON line per run (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch import nn  # noqa: E402


def to_torch_eager(model: nn.Module) -> nn.Module:
    """Re-bind the modules' forwards to their torch compositions (what the reference executes on any device)."""
    from vision_toolbox_b200.backbones.base import BaseBackbone
    from vision_toolbox_b200.components import ConvNormAct

    for m in model.modules():
        if isinstance(m, ConvNormAct):
            m.forward = types.MethodType(nn.Sequential.forward, m)
        elif hasattr(m, "_forward_cpu"):
            m.forward = types.MethodType(type(m)._forward_cpu, m)
        elif isinstance(m, BaseBackbone):
            m.get_feature_maps = types.MethodType(type(m)._features_cpu, m)
            m.forward = types.MethodType(lambda self, x: self.get_feature_maps(x)[-1], m)
    return model


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="cspdarknet53")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--res", type=int, default=176)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--eval", action="store_true", help="inference: get_feature_maps under no_grad (BASELINE config C5)")
    ap.add_argument("--no-channels-last", action="store_true")
    ap.add_argument("--no-sync-bn", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True

    from vision_toolbox_b200 import backbones
    from vision_toolbox_b200.parallel import split_decay_groups

    torch.manual_seed(0)

    # classifier.py:59-64: backbone -> AdaptiveAvgPool2d -> Flatten -> Linear
    class Net(nn.Module):
        def __init__(self, backbone, head):
            super().__init__()
            self.backbone, self.head = backbone, head

        def forward(self, x):
            return self.head(self.backbone(x).mean(dim=(2, 3)))

    backbone = getattr(backbones, args.model)()
    net = Net(backbone, nn.Linear(backbone.out_channels_list[-1], 1000)).to(dev)
    if not args.no_channels_last:
        net = net.to(memory_format=torch.channels_last)
    ddp = None
    if world > 1 and not args.eval:
        if not args.no_sync_bn:
            net = nn.SyncBatchNorm.convert_sync_batchnorm(net)   # configs/base.yaml:22
        to_torch_eager(net.backbone)
        ddp = nn.parallel.DistributedDataParallel(net, device_ids=[local])
    else:
        to_torch_eager(net.backbone)
    decay, no_decay = split_decay_groups([net])
    opt = torch.optim.SGD([{"params": decay, "weight_decay": 2e-5}, {"params": no_decay, "weight_decay": 0.0}],
                          lr=0.05, momentum=0.9, fused=True)

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    xs = [torch.rand(args.batch, 3, args.res, args.res, generator=g).to(dev) for _ in range(2)]
    if not args.no_channels_last:
        xs = [x.contiguous(memory_format=torch.channels_last) for x in xs]
    ys = [torch.randint(0, 1000, (args.batch,), generator=g).to(dev) for _ in range(2)]

    def train_step(x, y):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = (ddp if ddp is not None else net)(x)
            loss = F.cross_entropy(logits.float(), y, label_smoothing=0.1)
        loss.backward()
        opt.step()
        return loss

    def eval_step(x, y):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return net.backbone.get_feature_maps(x)[-1].float().mean()

    if args.eval:
        net.eval()
        step = eval_step
    else:
        net.train()
        step = train_step

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        out = step(xs[i % 2], ys[i % 2])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(xs[i % 2], ys[i % 2])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        print(json.dumps({
            "impl": "torch-eager-cudnn", "metric": "train images/sec" if not args.eval else "inference images/sec",
            "value": args.batch * world * args.steps / (ms / 1e3), "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "ms_per_step": ms / args.steps, "dtype": "bf16 autocast",
            "config": {"model": args.model, "batch_per_gpu": args.batch, "resolution": args.res,
                       "channels_last": not args.no_channels_last, "cudnn_benchmark": True,
                       "sync_bn": world > 1 and not args.no_sync_bn and not args.eval, "eval": args.eval},
            "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "last": float(out)}), flush=True)
    if world > 1:
        barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
