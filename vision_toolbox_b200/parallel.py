"""Data-parallel training step around a native backbone — the semantics the reference gets from Lightning
(``configs/base.yaml:16-23``: DDP gradient mean + SyncBatchNorm) and its trainer (``classifier.py:59-64``
head, ``:83-95`` step, ``:141-169`` SGD with weight decay only on conv/linear weights).

One process per GPU.  Two exchanges per step, both over NCCL/NVLink:
  * SyncBN: per BatchNorm layer the (sum, sum-of-squares) and (sum dz, sum dz*xhat) vectors are all-reduced
    inside the plan (engine.DistConfig) so every rank normalises with GLOBAL batch statistics;
  * gradients: all parameter gradients live in ONE flat fp32 buffer (parameter ``.grad`` are views), reduced
    in buckets and averaged.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .engine import DistConfig


def split_decay_groups(modules: list[nn.Module]):
    """classifier.py:141-169 — weight decay on conv / linear weights only, none on norm parameters and biases."""
    decay, no_decay = [], []
    for root in modules:
        for m in root.modules():
            for name, p in m.named_parameters(recurse=False):
                if isinstance(m, (nn.Conv2d, nn.Linear)) and name == "weight":
                    decay.append(p)
                else:
                    no_decay.append(p)
    return decay, no_decay


def bucket_ranges(sizes: list[int], bucket_elems: int) -> list[tuple[int, int]]:
    """Greedy contiguous buckets over a flat buffer: [(start, end)] in elements, each >= bucket_elems except the last."""
    out, start, acc = [], 0, 0
    for s in sizes:
        acc += s
        if acc - start >= bucket_elems:
            out.append((start, acc))
            start = acc
    if acc > start:
        out.append((start, acc))
    return out


class Trainer:
    def __init__(self, backbone: nn.Module, head: nn.Module, lr: float = 0.05, momentum: float = 0.9,
                 weight_decay: float = 2e-5, label_smoothing: float = 0.1, sync_bn: bool = True,
                 process_group=None, bucket_mb: float = 25.0):
        self.backbone, self.head = backbone, head
        self.label_smoothing = label_smoothing
        self.group = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size(process_group)
            # DDP broadcasts rank 0's parameters and buffers when it wraps the module
            for t in list(backbone.state_dict().values()) + list(head.state_dict().values()):
                dist.broadcast(t, src=0, group=process_group)
            backbone.__dict__["_vtb_dist"] = DistConfig(process_group, sync_bn=sync_bn)
        self.params = [p for p in list(backbone.parameters()) + list(head.parameters()) if p.requires_grad]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        off = 0
        for p, n in zip(self.params, sizes):
            p.grad = self.flat[off : off + n].view_as(p)
            off += n
        # the native backward writes backbone gradients straight into these views (engine.Runner grad_sink mode)
        backbone.__dict__["_vtb_grad_sink"] = True
        self.buckets = bucket_ranges(sizes, int(bucket_mb * 1024 * 1024 / 4))
        decay, no_decay = split_decay_groups([backbone, head])
        self.opt = torch.optim.SGD(
            [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}],
            lr=lr, momentum=momentum, fused=dev.type == "cuda")

    def forward_loss(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        f = self.backbone(x)  # (N, C, H, W) bf16
        pooled = f.float().mean(dim=(2, 3))  # AdaptiveAvgPool2d + Flatten (classifier.py:61-62)
        logits = self.head(pooled)
        return F.cross_entropy(logits, y, label_smoothing=self.label_smoothing)

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.flat.zero_()
        loss = self.forward_loss(x, y)
        loss.backward()
        if self.world > 1:
            import torch.distributed as dist

            works = [dist.all_reduce(self.flat[a:b], group=self.group, async_op=True) for a, b in self.buckets]
            for w in works:
                w.wait()
            self.flat.mul_(1.0 / self.world)
        self.opt.step()
        return loss.detach()
