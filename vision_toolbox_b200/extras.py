"""RandomMixup / RandomCutmix / RandomCutMixMixUp - the batch transforms of the reference trainer (extras.py:14-109,
used at classifier.py:66-67, 86-87), same constructors and ``forward(batch, target) -> (batch, target)``.

CPU tensors: the reference's semantics with the reference's sequence of host RNG draws (same seed -> same result).
CUDA tensors: the decision (apply?, mixup or cutmix), lambda and the cut box are SAMPLED ON THE DEVICE and consumed by one
kernel (``vtb_mix_images``) - no ``.item()`` / ``float(tensor)`` host synchronisation anywhere, so the transform can sit
inside a captured training step.  Targets become probability vectors exactly as in the reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor, nn

__all__ = ["RandomMixup", "RandomCutmix", "RandomCutMixMixUp", "mix_apply"]

MODE_NONE, MODE_MIXUP, MODE_CUTMIX = 0, 1, 2


def _soft_targets(target: Tensor, num_classes: int, dtype: torch.dtype) -> Tensor:
    if target.ndim == 1:
        return F.one_hot(target, num_classes=num_classes).to(dtype=dtype)
    return target.clone()


def mix_apply(batch: Tensor, target: Tensor, params: Tensor, lam_target: Tensor) -> tuple[Tensor, Tensor]:
    """Apply a sampled transform.  params: six floats {mode, lambda, x1, y1, x2, y2} on batch.device; lam_target: two
    weights {w, 1 - w} of the un-rolled / rolled targets, each rounded from the double the reference computes (w = lambda
    for mixup, 1 - box area fraction for cutmix, 1 when nothing is applied)."""
    if batch.is_cuda:
        from . import _lib
        from ._lib import check

        x = batch if (batch.dtype == torch.float32 and batch.is_contiguous()) else batch.float().contiguous()
        out = torch.empty_like(x)
        n, c, h, w = x.shape
        check(_lib.lib().vtb_mix_images(x.data_ptr(), out.data_ptr(), n, c, h, w, params.data_ptr(),
                                        torch.cuda.current_stream(x.device).cuda_stream), "vtb_mix_images")
        out = out.to(batch.dtype)
    else:
        mode, lam = int(params[0]), params[1]
        rolled = batch.roll(1, 0)
        if mode == MODE_MIXUP:
            out = batch * lam + rolled * (1.0 - lam)
        elif mode == MODE_CUTMIX:
            x1, y1, x2, y2 = (int(v) for v in params[2:6])
            out = batch.clone()
            out[:, :, y1:y2, x1:x2] = rolled[:, :, y1:y2, x1:x2]
        else:
            out = batch.clone()
    lt = lam_target.to(target.dtype)
    return out, target * lt[0] + target.roll(1, 0) * lt[1]


class _Mix(nn.Module):
    mode = MODE_NONE

    def __init__(self, num_classes: int, p: float = 0.5, alpha: float = 1, inplace: bool = False):
        super().__init__()
        self.num_classes, self.p, self.alpha, self.inplace = num_classes, p, alpha, inplace

    # -- host sampling: the reference's draws, in the reference's order (extras.py:29-36 / :62-84)
    def _sample_host(self, h: int, w: int):
        if torch.rand(1).item() >= self.p:
            return None
        lam = float(torch._sample_dirichlet(torch.tensor([float(self.alpha), float(self.alpha)]))[0])
        if self.mode == MODE_MIXUP:
            return (torch.tensor([MODE_MIXUP, lam, 0, 0, 0, 0], dtype=torch.float32),
                    torch.tensor([lam, 1.0 - lam], dtype=torch.float32))
        r_x, r_y = torch.randint(w, (1,)), torch.randint(h, (1,))
        r = 0.5 * math.sqrt(1.0 - lam)
        rw, rh = int(r * w), int(r * h)
        x1, y1 = int(torch.clamp(r_x - rw, min=0)), int(torch.clamp(r_y - rh, min=0))
        x2, y2 = int(torch.clamp(r_x + rw, max=w)), int(torch.clamp(r_y + rh, max=h))
        lam2 = float(1.0 - (x2 - x1) * (y2 - y1) / (w * h))
        return (torch.tensor([MODE_CUTMIX, lam, x1, y1, x2, y2], dtype=torch.float32),
                torch.tensor([lam2, 1.0 - lam2], dtype=torch.float32))

    # -- device sampling: same distributions, no host round trip
    def _sample_device(self, h: int, w: int, dev: torch.device):
        apply = torch.rand(1, device=dev) < self.p
        conc = torch.full((2,), float(self.alpha), device=dev)
        lam = torch._sample_dirichlet(conc)[0]
        if self.mode == MODE_MIXUP:
            box = torch.zeros(4, device=dev)
            lam_t = lam.double()
        else:
            r_x = torch.randint(w, (1,), device=dev).float()
            r_y = torch.randint(h, (1,), device=dev).float()
            r = 0.5 * torch.sqrt(1.0 - lam.double())
            rw, rh = torch.floor(r * w).float(), torch.floor(r * h).float()
            x1, y1 = (r_x - rw).clamp(min=0), (r_y - rh).clamp(min=0)
            x2, y2 = (r_x + rw).clamp(max=w), (r_y + rh).clamp(max=h)
            box = torch.cat([x1, y1, x2, y2])
            lam_t = (1.0 - ((x2 - x1) * (y2 - y1)).double() / float(w * h))[0]
        mode = torch.where(apply, torch.full((1,), float(self.mode), device=dev), torch.zeros(1, device=dev))
        params = torch.cat([mode, lam.reshape(1).float(), box.float()])
        lam_t = torch.where(apply[0], lam_t, torch.ones((), device=dev, dtype=torch.float64))
        return params, torch.stack([lam_t, 1.0 - lam_t]).float()

    def forward(self, batch: Tensor, target: Tensor) -> tuple[Tensor, Tensor]:
        if batch.ndim != 4:
            raise ValueError(f"Batch ndim should be 4. Got {batch.ndim}")
        target = _soft_targets(target, self.num_classes, batch.dtype)
        h, w = batch.shape[-2:]
        if batch.is_cuda:
            params, lam_t = self._sample_device(h, w, batch.device)
            return mix_apply(batch, target, params, lam_t)
        sampled = self._sample_host(h, w)
        if sampled is None:
            return batch.clone(), target
        return mix_apply(batch, target, *sampled)


class RandomMixup(_Mix):
    mode = MODE_MIXUP


class RandomCutmix(_Mix):
    mode = MODE_CUTMIX


class RandomCutMixMixUp(nn.Module):
    def __init__(self, num_classes: int, cutmix_alpha: float, mixup_alpha: float, inplace: bool = False):
        super().__init__()
        if cutmix_alpha == 0 and mixup_alpha == 0:
            raise ValueError
        self.cutmix = RandomCutmix(num_classes, p=1, alpha=cutmix_alpha, inplace=inplace) if cutmix_alpha > 0 else None
        self.mixup = RandomMixup(num_classes, p=1, alpha=mixup_alpha, inplace=inplace) if mixup_alpha > 0 else None

    def forward(self, batch: Tensor, target: Tensor) -> tuple[Tensor, Tensor]:
        if not batch.is_cuda:
            if self.cutmix is None or torch.rand(1).item() >= 0.5:
                return self.mixup(batch, target)
            return self.cutmix(batch, target)
        # device: draw both candidates' parameters and select with a device-side coin (no branch on the host)
        h, w = batch.shape[-2:]
        dev = batch.device
        tgt = _soft_targets(target, self.mixup.num_classes if self.mixup is not None else self.cutmix.num_classes, batch.dtype)
        if self.cutmix is None:
            return mix_apply(batch, tgt, *self.mixup._sample_device(h, w, dev))
        if self.mixup is None:
            # reference quirk (extras.py:103-105): with mixup_alpha == 0 the mixup branch would fail; cutmix is what can run
            return mix_apply(batch, tgt, *self.cutmix._sample_device(h, w, dev))
        pm, lm = self.mixup._sample_device(h, w, dev)
        pc, lc = self.cutmix._sample_device(h, w, dev)
        use_mix = torch.rand(1, device=dev) >= 0.5
        return mix_apply(batch, tgt, torch.where(use_mix, pm, pc), torch.where(use_mix, lm, lc))
