"""Darknet / CSPDarknet / Darknet-YOLOv5 — drop-in for reference vision_toolbox/backbones/darknet.py.

Module tree, constructor signatures, attribute names and therefore state_dict keys are the reference's
(darknet.py:20-137).  Each block additionally knows how to emit itself into the native planner:
residual adds ride on the normalise+ReLU pass of the block's last ConvNormAct (darknet.py:28) and the CSP
concatenation (darknet.py:53) is two channel slices of one buffer.
"""
from __future__ import annotations

from typing import Callable, NamedTuple

import torch
from torch import Tensor, nn

from ..components import ConvNormAct
from .base import BaseBackbone

__all__ = [
    "Darknet", "DarknetYOLOv5", "DarknetBlock", "DarknetStage", "CSPDarknetStage",
    "darknet19", "darknet53", "cspdarknet53",
    "darknet_yolov5n", "darknet_yolov5s", "darknet_yolov5m", "darknet_yolov5l", "darknet_yolov5x",
]

_BASE_URL = "https://github.com/gau-nernst/vision-toolbox/releases/download/v0.0.1/"


class _NativeMixin:
    """forward(): CUDA tensors go through the planner, CPU tensors through the torch composition."""

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            from ..engine import run_native

            return run_native(self, x)[0]
        return self._forward_cpu(x)


class DarknetBlock(_NativeMixin, nn.Module):
    def __init__(self, in_channels: int, expansion: float = 0.5) -> None:
        super().__init__()
        mid_channels = int(in_channels * expansion)
        self.conv1 = ConvNormAct(in_channels, mid_channels, 1)
        self.conv2 = ConvNormAct(mid_channels, in_channels)

    def _forward_cpu(self, x: Tensor) -> Tensor:
        return x + self.conv2(self.conv1(x))

    def _emit(self, g, x, out=None):
        return self.conv2._emit(g, self.conv1._emit(g, x), residual=x, out=out)


def _emit_blocks(blocks: nn.Sequential, g, x, out=None):
    n = len(blocks)
    for i, blk in enumerate(blocks):
        x = blk._emit(g, x, out=out if i == n - 1 else None)
    return x


class DarknetStage(_NativeMixin, nn.Sequential):
    def __init__(self, n: int, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.conv = ConvNormAct(in_channels, out_channels, stride=2)
        self.blocks = nn.Sequential(*[DarknetBlock(out_channels) for _ in range(n)])

    def _forward_cpu(self, x: Tensor) -> Tensor:
        return self.blocks(self.conv(x))

    def _emit(self, g, x):
        return _emit_blocks(self.blocks, g, self.conv._emit(g, x))


class CSPDarknetStage(_NativeMixin, nn.Module):
    def __init__(self, n: int, in_channels: int, out_channels: int) -> None:
        assert n > 0
        super().__init__()
        self.conv = ConvNormAct(in_channels, out_channels, stride=2)

        half_channels = out_channels // 2
        self.conv1 = ConvNormAct(out_channels, half_channels, 1)
        self.conv2 = ConvNormAct(out_channels, half_channels, 1)
        self.blocks = nn.Sequential(*[DarknetBlock(half_channels, expansion=1) for _ in range(n)])
        self.out_conv = ConvNormAct(out_channels, out_channels, 1)

    def _forward_cpu(self, x: Tensor) -> Tensor:
        out = self.conv(x)
        out = torch.cat([self.conv1(out), self.blocks(self.conv2(out))], dim=1)
        return self.out_conv(out)

    def _emit(self, g, x):
        out = self.conv._emit(g, x)
        half = self.conv1.conv.out_channels
        cat = g.new_buffer(out.n, out.h, out.w, 2 * half)
        # torch.cat([conv1(out), blocks(conv2(out))], 1): both producers write their channel slice directly;
        # conv1 and conv2 read the same tensor -> one side-by-side convolution in training plans (engine.Graph)
        _, b = g.conv_norm_act_pair(self.conv1, self.conv2, out, out_a=g.slice(cat, 0, half))
        _emit_blocks(self.blocks, g, b, out=g.slice(cat, half, half))
        return self.out_conv._emit(g, g.slice(cat, 0, 2 * half))


class DarknetStageConfig(NamedTuple):
    n_blocks: int
    out_channels: int


class Darknet(BaseBackbone):
    def __init__(
        self,
        stem_channels: int,
        stage_configs: list[DarknetStageConfig | tuple[int, int]],
        stage_cls: Callable[..., nn.Module] = DarknetStage,
    ):
        assert len(stage_configs) > 0
        super().__init__()
        self.out_channels_list = tuple(cfg[1] for cfg in stage_configs)
        self.stride = 32

        self.stem = ConvNormAct(3, stem_channels)
        self.stages = nn.ModuleList()
        in_ch = stem_channels
        for n_blocks, out_ch in stage_configs:
            stage = stage_cls(n_blocks, in_ch, out_ch) if n_blocks else ConvNormAct(in_ch, out_ch, 3, 2)
            self.stages.append(stage)
            in_ch = out_ch

    def _features_cpu(self, x: Tensor) -> list[Tensor]:
        outputs = [self.stem(x)]
        for s in self.stages:
            outputs.append(s(outputs[-1]))
        return outputs[1:]

    def _emit(self, g, x):
        outputs = [self.stem._emit(g, x)]
        for s in self.stages:
            outputs.append(s._emit(g, outputs[-1]))
        return outputs[1:]

    @staticmethod
    def from_config(variant: str, pretrained: bool = False) -> "Darknet":
        n_blocks_list, stage_cls, ckpt = dict(
            darknet19=((0, 1, 1, 2, 2), DarknetStage, "darknet19-2cb641ca.pth"),
            darknet53=((1, 2, 8, 8, 4), DarknetStage, "darknet53-94427f5b.pth"),
            cspdarknet53=((1, 2, 8, 8, 4), CSPDarknetStage, "cspdarknet53-3bfa0423.pth"),
        )[variant]
        stage_configs = list(zip(n_blocks_list, (64, 128, 256, 512, 1024)))
        m = Darknet(32, stage_configs, stage_cls)
        if pretrained:
            m._load_state_dict_from_url(_BASE_URL + ckpt)
        return m


class DarknetYOLOv5(BaseBackbone):
    def __init__(self, stem_channels: int, stage_configs: list[DarknetStageConfig | tuple[int, int]]) -> None:
        super().__init__()
        self.out_channels_list = (stem_channels,) + tuple(cfg[1] for cfg in stage_configs)
        self.stride = 2 ** len(self.out_channels_list)

        self.stem = ConvNormAct(3, stem_channels, 6, 2)
        self.stages = nn.ModuleList()
        in_ch = stem_channels
        for n_blocks, out_ch in stage_configs:
            self.stages.append(CSPDarknetStage(n_blocks, in_ch, out_ch))
            in_ch = out_ch

    def _features_cpu(self, x: Tensor) -> list[Tensor]:
        outputs = [self.stem(x)]
        for s in self.stages:
            outputs.append(s(outputs[-1]))
        return outputs

    def _emit(self, g, x):
        outputs = [self.stem._emit(g, x)]
        for s in self.stages:
            outputs.append(s._emit(g, outputs[-1]))
        return outputs

    @staticmethod
    def from_config(variant: str, pretrained: bool = False) -> "DarknetYOLOv5":
        depth_scale, width_scale, ckpt = dict(
            n=(1 / 3, 1 / 4, "darknet_yolov5n-68f182f1.pth"),
            s=(1 / 3, 1 / 2, "darknet_yolov5s-175f7462.pth"),
            m=(2 / 3, 3 / 4, "darknet_yolov5m-9866aa40.pth"),
            l=(1 / 1, 1 / 1, "darknet_yolov5l-8e25d388.pth"),
            x=(4 / 3, 5 / 4, "darknet_yolov5x-0ed0c035.pth"),
        )[variant]
        stage_configs = [
            (int(d * depth_scale), int(w * width_scale)) for d, w in zip((3, 6, 9, 3), (128, 256, 512, 1024))
        ]
        m = DarknetYOLOv5(int(64 * width_scale), stage_configs)
        if pretrained:
            m._load_state_dict_from_url(_BASE_URL + ckpt)
        return m


# factory functions named after the released checkpoints (reference README.md:25-31)
def darknet19(pretrained: bool = False) -> Darknet:
    return Darknet.from_config("darknet19", pretrained)


def darknet53(pretrained: bool = False) -> Darknet:
    return Darknet.from_config("darknet53", pretrained)


def cspdarknet53(pretrained: bool = False) -> Darknet:
    return Darknet.from_config("cspdarknet53", pretrained)


def darknet_yolov5n(pretrained: bool = False) -> DarknetYOLOv5:
    return DarknetYOLOv5.from_config("n", pretrained)


def darknet_yolov5s(pretrained: bool = False) -> DarknetYOLOv5:
    return DarknetYOLOv5.from_config("s", pretrained)


def darknet_yolov5m(pretrained: bool = False) -> DarknetYOLOv5:
    return DarknetYOLOv5.from_config("m", pretrained)


def darknet_yolov5l(pretrained: bool = False) -> DarknetYOLOv5:
    return DarknetYOLOv5.from_config("l", pretrained)


def darknet_yolov5x(pretrained: bool = False) -> DarknetYOLOv5:
    return DarknetYOLOv5.from_config("x", pretrained)
