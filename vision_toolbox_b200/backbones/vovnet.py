"""VoVNet (V1 / V2-eSE) — drop-in for reference vision_toolbox/backbones/vovnet.py.

Same module tree and state_dict keys (vovnet.py:20-136).  In the native plan the one-shot-aggregation
concatenation (vovnet.py:51-55) is a single NHWC buffer: the block input is *re-homed* into channel slice 0
by its producer and each 3x3 ConvNormAct writes its own slice, so no concat kernel runs; the eSE gate and the
identity add (vovnet.py:58-61) are one pass.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
from torch import Tensor, nn

from ..components import ConvNormAct
from .base import BaseBackbone
from .darknet import _NativeMixin

__all__ = [
    "VoVNet", "OSABlock", "ESEBlock",
    "vovnet27_slim", "vovnet39", "vovnet57",
    "vovnet19_slim_ese", "vovnet19_ese", "vovnet39_ese", "vovnet57_ese", "vovnet99_ese",
]

_BASE_URL = "https://github.com/gau-nernst/vision-toolbox/releases/download/v0.0.1/"


class ESEBlock(nn.Module):
    def __init__(self, num_channels: int) -> None:
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.linear = nn.Conv2d(num_channels, num_channels, 1)
        self.gate = nn.Hardsigmoid(inplace=True)

    def forward(self, x: Tensor) -> Tensor:
        # only reached with CPU tensors (OSABlock emits the gate natively for CUDA tensors)
        return x * self.gate(self.linear(self.pool(x)))


class OSABlock(_NativeMixin, nn.Module):
    def __init__(
        self,
        in_channels: int,
        mid_channels: int,
        num_layers: int,
        out_channels: int,
        ese: bool = True,
    ) -> None:
        super().__init__()
        self.convs = nn.ModuleList(
            [ConvNormAct(in_channels if i == 0 else mid_channels, mid_channels) for i in range(num_layers)]
        )
        concat_channels = in_channels + mid_channels * num_layers
        self.out_conv = ConvNormAct(concat_channels, out_channels, 1)

        self.ese = ESEBlock(out_channels) if ese else None
        self.residual = in_channels == out_channels

    def _forward_cpu(self, x: Tensor) -> Tensor:
        outputs = [x]
        for conv in self.convs:
            outputs.append(conv(outputs[-1]))
        out = self.out_conv(torch.cat(outputs, dim=1))
        if self.ese is not None:
            out = self.ese(out)
        if self.residual:
            out = out + x
        return out

    def _emit(self, g, x):
        in_c = self.convs[0].conv.in_channels
        mid = self.convs[0].conv.out_channels
        n_layers = len(self.convs)
        have = g.input_c if (x.is_input and x is g.input) else x.c
        if have != in_c:
            raise RuntimeError(f"Given groups=1, weight of size {list(self.convs[0].conv.weight.shape)}, expected input "
                               f"to have {in_c} channels, but got {have} channels instead")
        cat = g.new_buffer(x.n, x.h, x.w, in_c + mid * n_layers)
        if not g.rehome(x, cat, 0):
            # the block is called on its own (reference vovnet.py:50: OSABlock.forward(x)): x is the plan's input and
            # cannot move - copy it into slice 0 (its gradient flows back through the copy)
            x = g.copy_into(x, cat, 0)
        cur = x
        for i, conv in enumerate(self.convs):
            cur = conv._emit(g, cur, out=g.slice(cat, in_c + i * mid, mid))
        full = g.slice(cat, 0, cat.c)
        res = x if self.residual else None
        if self.ese is None:
            return self.out_conv._emit(g, full, residual=res)
        return g.ese(self.ese, self.out_conv._emit(g, full), residual=res)


class VoVNetStageConfig(NamedTuple):
    n_blocks: int
    mid_channels: int
    n_layers: int
    out_channels: int


class _Stage(_NativeMixin, nn.Sequential):
    """max_pool -> module_0 -> module_1 ... (child names fixed by reference vovnet.py:93-98)."""

    def _forward_cpu(self, x: Tensor) -> Tensor:
        return nn.Sequential.forward(self, x)

    def _emit(self, g, x):
        for name, child in self.named_children():
            x = g.maxpool(x) if name == "max_pool" else child._emit(g, x)
        return x


class VoVNet(BaseBackbone):
    def __init__(
        self,
        stem_channels: int,
        stage_configs: list[VoVNetStageConfig | tuple[int, int, int, int]],
        ese: bool = True,
    ) -> None:
        super().__init__()
        self.out_channels_list = (stem_channels,) + tuple(cfg[3] for cfg in stage_configs)
        self.stride = 2 ** len(self.out_channels_list)

        self.stem = nn.Sequential(
            ConvNormAct(3, stem_channels // 2, 3, 2),
            ConvNormAct(stem_channels // 2, stem_channels // 2),
            ConvNormAct(stem_channels // 2, stem_channels),
        )

        self.stages = nn.ModuleList()
        in_ch = stem_channels
        for n_blocks, mid_ch, n_layers, out_ch in stage_configs:
            stage = _Stage()
            stage.add_module("max_pool", nn.MaxPool2d(3, 2, 1))
            for i in range(n_blocks):
                stage.add_module(f"module_{i}", OSABlock(in_ch, mid_ch, n_layers, out_ch, ese))
                in_ch = out_ch
            self.stages.append(stage)

    def _features_cpu(self, x: Tensor) -> list[Tensor]:
        outputs = [self.stem(x)]
        for s in self.stages:
            outputs.append(s(outputs[-1]))
        return outputs

    def _emit(self, g, x):
        for cna in self.stem:
            x = cna._emit(g, x)
        outputs = [x]
        for s in self.stages:
            outputs.append(s._emit(g, outputs[-1]))
        return outputs

    @staticmethod
    def from_config(variant: int, slim: bool = False, ese: bool = False, pretrained: bool = False) -> "VoVNet":
        stem_channels = 128
        mid_channels_list = (64, 80, 96, 112) if slim else (128, 160, 192, 224)
        out_channels_list = (128, 256, 384, 512) if slim else (256, 512, 768, 1024)
        n_blocks_list, n_layers_list = {
            19: ((1, 1, 1, 1), (3, 3, 3, 3)),
            27: ((1, 1, 1, 1), (5, 5, 5, 5)),
            39: ((1, 1, 2, 2), (5, 5, 5, 5)),
            57: ((1, 1, 4, 3), (5, 5, 5, 5)),
            99: ((1, 3, 9, 3), (5, 5, 5, 5)),
        }[variant]
        stage_configs = list(zip(n_blocks_list, mid_channels_list, n_layers_list, out_channels_list))
        m = VoVNet(stem_channels, stage_configs, ese)

        if pretrained:
            ckpt = {
                (27, True, False): "vovnet27_slim-dd43306a.pth",
                (39, False, False): "vovnet39-4c79d629.pth",
                (57, False, False): "vovnet57-ecb9cc34.pth",
                (19, True, True): "vovnet19_slim_ese-f8075640.pth",
                (19, False, True): "vovnet19_ese-a077657e.pth",
                (39, False, True): "vovnet39_ese-9ce81b0d.pth",
                (57, False, True): "vovnet57_ese-ae1a7f89.pth",
                (99, False, True): "vovnet99_ese-713f3062.pth",
            }[(variant, slim, ese)]
            m._load_state_dict_from_url(_BASE_URL + ckpt)

        return m


def vovnet27_slim(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(27, slim=True, ese=False, pretrained=pretrained)


def vovnet39(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(39, pretrained=pretrained)


def vovnet57(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(57, pretrained=pretrained)


def vovnet19_slim_ese(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(19, slim=True, ese=True, pretrained=pretrained)


def vovnet19_ese(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(19, ese=True, pretrained=pretrained)


def vovnet39_ese(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(39, ese=True, pretrained=pretrained)


def vovnet57_ese(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(57, ese=True, pretrained=pretrained)


def vovnet99_ese(pretrained: bool = False) -> VoVNet:
    return VoVNet.from_config(99, ese=True, pretrained=pretrained)
