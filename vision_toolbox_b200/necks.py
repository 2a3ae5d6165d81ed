"""FPN / PAN - drop-in for reference vision_toolbox/necks.py:45-120 (BiFPN, necks.py:123+, is outside the conv-BN-ReLU
hot path: SURVEY.md section 8f.2 names FPN and PAN).

Same constructors, attribute names and therefore state_dict keys (``lateral_convs.{i}.{weight,bias}``,
``output_convs.{i}.conv.weight`` ...; PAN: ``top_down.*`` / ``bottom_up.*``).  CPU tensors run the reference's torch
composition.  CUDA tensors (the NHWC bf16 feature maps a native backbone returns, or anything convertible) run on the
sm_100a kernels: the biased 1x1 lateral convolutions through the planner's bias unit (one launch, bias in the epilogue),
nearest x2 / x0.5 resize + "sum" fuse through ONE kernel (``vtb_resize2_add``; "concat" writes the two halves of a channel
concatenation in place), output blocks through the native ConvNormAct.
"""
from __future__ import annotations

from typing import Callable

import torch
from torch import Tensor, nn

from .components import ConvNormAct

__all__ = ["FPN", "PAN"]


def aggregate_concat(x: list[Tensor]) -> Tensor:
    return torch.cat(x, dim=1)


def aggregate_sum(x: list[Tensor]) -> Tensor:
    out = x[0]
    for o in x[1:]:
        out = out + o
    return out


def aggregate_avg(x: list[Tensor]) -> Tensor:
    return aggregate_sum(x) / len(x)


def aggregate_max(x: list[Tensor]) -> Tensor:
    out = x[0]
    for o in x[1:]:
        out = torch.maximum(out, o)
    return out


_aggregate_functions = {"concat": aggregate_concat, "sum": aggregate_sum, "avg": aggregate_avg, "max": aggregate_max}


class _BiasConvUnit(nn.Module):
    """Planner face of a bare ``nn.Conv2d`` with bias (not registered in the neck: state_dict keys stay the reference's)."""

    def __init__(self, conv: nn.Conv2d):
        super().__init__()
        self.conv = conv

    def _emit(self, g, x):
        return g.conv_bias(self, x)

    def forward(self, x: Tensor) -> Tensor:
        from .engine import run_native

        return run_native(self, x)[0]


def _nhwc_bf16(t: Tensor) -> Tensor:
    """(N, C, H, W) bf16 tensor with channels-last strides (pointer + pixel pitch for the kernels); no copy if it is one."""
    if (t.dtype == torch.bfloat16 and t.stride(1) == 1 and t.stride(3) % 8 == 0 and t.stride(2) == t.shape[3] * t.stride(3)
            and t.stride(0) == t.shape[2] * t.shape[3] * t.stride(3) and t.data_ptr() % 16 == 0):
        return t
    return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


class _ResizeFuse(torch.autograd.Function):
    """fuse([a, resize(b)]) of necks.py:71 / :77 with nearest x2 (up) or x0.5 (down): 'sum' -> a + resize(b); 'concat' ->
    cat([a, resize(b)], 1).  One kernel for the sum; the concatenation is written in place (two launches, no cat)."""

    @staticmethod
    def forward(ctx, a: Tensor, b: Tensor, up: bool, concat: bool):
        from . import _lib
        from ._lib import check

        L = _lib.lib()
        a, b = _nhwc_bf16(a), _nhwc_bf16(b)
        n, c, h, w = a.shape
        nb, cb, hb, wb = b.shape
        if (n, c) != (nb, cb) or (up and (h, w) != (2 * hb, 2 * wb)) or (not up and (h, w) != (hb // 2, wb // 2)):
            # what the reference's `+` / torch.cat raises on mismatching pyramid levels
            raise RuntimeError(f"The size of tensor a ({tuple(a.shape)}) must match the size of the resized tensor b "
                               f"({tuple(b.shape)}, scale {'2' if up else '0.5'})")
        if c % 8:
            raise NotImplementedError("native necks need a channel count that is a multiple of 8")
        st = torch.cuda.current_stream(a.device).cuda_stream
        co = 2 * c if concat else c
        out = torch.empty((n, co, h, w), dtype=torch.bfloat16, device=a.device, memory_format=torch.channels_last)
        if concat:
            check(L.vtb_grad_add(out.data_ptr(), co, a.data_ptr(), a.stride(3), n * h * w, c, 0, st), "vtb_grad_add(concat)")
            check(L.vtb_resize2_add(0, 0, b.data_ptr(), b.stride(3), n, h, w, c, hb, wb, int(up),
                                    out.data_ptr() + 2 * c, co, st), "vtb_resize2_add(concat)")
        else:
            check(L.vtb_resize2_add(a.data_ptr(), a.stride(3), b.data_ptr(), b.stride(3), n, h, w, c, hb, wb, int(up),
                                    out.data_ptr(), co, st), "vtb_resize2_add")
        ctx.geom = (n, c, h, w, hb, wb, up, concat)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout: Tensor):
        from . import _lib
        from ._lib import check

        L = _lib.lib()
        n, c, h, w, hb, wb, up, concat = ctx.geom
        gout = _nhwc_bf16(gout)
        st = torch.cuda.current_stream(gout.device).cuda_stream
        ld = gout.stride(3)
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = gout[:, :c] if concat else gout      # identity (a channel slice of the concatenation)
        if ctx.needs_input_grad[1]:
            gb = torch.empty((n, c, hb, wb), dtype=torch.bfloat16, device=gout.device, memory_format=torch.channels_last)
            src = gout.data_ptr() + (2 * c if concat else 0)
            check(L.vtb_resize2_add_bwd(src, ld, n, h, w, c, gb.data_ptr(), c, hb, wb, int(up), 0, st), "vtb_resize2_add_bwd")
        return ga, gb, None, None


# https://arxiv.org/abs/1612.03144
class FPN(nn.Module):
    def __init__(
        self,
        in_channels_list: list[int],
        out_channels: int = 256,
        fuse_fn: str = "sum",
        block: Callable[[int, int], nn.Module] = ConvNormAct,
        interpolation_mode: str = "nearest",
        top_down: bool = True,
    ):
        super().__init__()
        self.fuse = _aggregate_functions[fuse_fn]
        self.out_channels = out_channels
        self.top_down = top_down

        self.lateral_convs = nn.ModuleList(
            [
                nn.Conv2d(in_c, out_channels, kernel_size=1) if in_c != out_channels else nn.Identity()
                for in_c in in_channels_list
            ]
        )
        self.upsample = nn.Upsample(scale_factor=2.0 if top_down else 0.5, mode=interpolation_mode)
        in_c = out_channels if fuse_fn == "sum" else out_channels * 2
        self.output_convs = nn.ModuleList([block(in_c, out_channels) for _ in range(len(in_channels_list) - 1)])
        self._fuse_name, self._mode = fuse_fn, interpolation_mode

    # ---- native pieces ------------------------------------------------------------------------------------------
    def _lateral(self, i: int, x: Tensor) -> Tensor:
        conv = self.lateral_convs[i]
        if isinstance(conv, nn.Identity):
            return x
        units = self.__dict__.setdefault("_vtb_units", {})
        unit = units.get(i)
        if unit is None or unit.conv is not conv:
            unit = units[i] = _BiasConvUnit(conv)
        unit.train(self.training)
        return unit(x)

    def _fuse_native(self, a: Tensor, b: Tensor) -> Tensor:
        if self._fuse_name not in ("sum", "concat") or self._mode != "nearest":
            raise NotImplementedError("native necks implement nearest resize with the 'sum' and 'concat' aggregates; "
                                      "other options run on CPU tensors only")
        return _ResizeFuse.apply(a, b, self.top_down, self._fuse_name == "concat")

    def _fuse_top_down(self, x: list[Tensor], fuse) -> list[Tensor]:
        for i, output_conv in enumerate(self.output_convs):
            x[-2 - i] = fuse(x[-2 - i], x[-1 - i])  # 2, 1, 0
            x[-2 - i] = output_conv(x[-2 - i])
        return x

    def _fuse_bottom_up(self, x: list[Tensor], fuse) -> list[Tensor]:
        for i, output_conv in enumerate(self.output_convs):
            x[i + 1] = fuse(x[i + 1], x[i])  # 1, 2, 3
            x[i + 1] = output_conv(x[i + 1])
        return x

    # input feature maps are ordered from bottom (largest) to top (smallest)
    def forward(self, x: list[Tensor]) -> list[Tensor]:
        assert len(x) == len(self.lateral_convs)
        if x[0].is_cuda:
            from .engine import get_precision

            if get_precision() == "fp32":
                raise NotImplementedError("native necks run in bf16 mode only (no fp32 parity kernels for the resize fuse)")
            outputs = [self._lateral(i, x[i]) for i in range(len(x))]
            fuse = self._fuse_native
        else:
            outputs = [l_conv(x[i]) for i, l_conv in enumerate(self.lateral_convs)]
            fuse = lambda a, b: self.fuse([a, self.upsample(b)])
        if self.top_down:
            return self._fuse_top_down(outputs, fuse)
        return self._fuse_bottom_up(outputs, fuse)


# https://arxiv.org/abs/1803.01534
class PAN(nn.Module):
    def __init__(
        self,
        in_channels_list: list[int],
        out_channels: int = 256,
        fuse_fn: str = "sum",
        block: Callable[[int, int], nn.Module] = ConvNormAct,
        interpolation_mode: str = "nearest",
    ):
        super().__init__()
        self.top_down = FPN(
            in_channels_list,
            out_channels,
            fuse_fn=fuse_fn,
            block=block,
            interpolation_mode=interpolation_mode,
        )
        self.bottom_up = FPN(
            [out_channels] * len(in_channels_list),
            out_channels,
            fuse_fn=fuse_fn,
            block=block,
            interpolation_mode=interpolation_mode,
        )

    def forward(self, x: list[Tensor]) -> list[Tensor]:
        outputs = self.top_down(x)
        outputs = self.bottom_up(outputs)
        return outputs
