"""ConvNormAct — drop-in for reference vision_toolbox/components.py:13-46.

Same constructor, same child names (``conv``, ``norm``, ``act`` → identical state_dict keys), same
initialisation.  CUDA tensors are executed by the sm_100a kernels behind include/vtb.h (implicit-GEMM
convolution with the BatchNorm statistics reduced in its epilogue, then one normalise+ReLU(+residual) pass);
CPU tensors run the plain ``nn.Sequential`` composition so the reference's own CPU tests (shape checks,
``torch.jit.trace``, tests/test_backbones.py:39-78) keep working.
"""
from __future__ import annotations

import math
from functools import partial

import torch
from torch import Tensor, nn

__all__ = ["ConvNormAct", "ConvBnAct"]


class ConvNormAct(nn.Sequential):
    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        kernel_size: int = 3,
        stride: int = 1,
        dilation: int = 1,
        groups: int = 1,
        norm: str = "bn",
        act: str = "relu",
    ):
        super().__init__()
        # reference components.py:26-35 — padding = ceil((k - s) / 2), bias only when there is no norm
        self.conv = nn.Conv2d(
            in_channels,
            out_channels,
            kernel_size,
            stride=stride,
            padding=math.ceil((kernel_size - stride) / 2),
            dilation=dilation,
            groups=groups,
            bias=norm == "none",
        )
        self.norm = dict(none=nn.Identity, bn=nn.BatchNorm2d)[norm](out_channels)
        self.act = dict(
            none=nn.Identity,
            relu=partial(nn.ReLU, True),
            leaky_relu=partial(nn.LeakyReLU, 0.2, True),
            swish=partial(nn.SiLU, True),
            silu=partial(nn.SiLU, True),
            gelu=nn.GELU,
        )[act]()
        # reference components.py:45-46
        if act in ("relu", "leaky_relu"):
            nn.init.kaiming_normal_(self.conv.weight, 0.2, "fan_out", act)

    # ---- native (sm_100a) path -------------------------------------------------------------
    def native_supported(self) -> bool:
        """True for the configurations the Darknet / VoVNet hot path uses."""
        c = self.conv
        return (
            isinstance(self.norm, nn.BatchNorm2d)
            and isinstance(self.act, (nn.ReLU, nn.Identity))
            and c.groups == 1
            and c.dilation == (1, 1)
            and c.kernel_size[0] == c.kernel_size[1]
            and c.stride[0] == c.stride[1]
            and c.stride[0] in (1, 2)
            and c.bias is None
            and c.out_channels % 16 == 0
            and (c.in_channels % 16 == 0 or c.in_channels <= 16)
            and c.kernel_size[0] ** 2 <= 36
        )

    def act_is_relu(self) -> bool:
        return isinstance(self.act, nn.ReLU)

    def _emit(self, g, x, residual=None, out=None):
        return g.conv_norm_act(self, x, residual=residual, out=out)

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda and self.native_supported():
            from .engine import run_native

            return run_native(self, x)[0]
        if x.is_cuda:
            raise NotImplementedError(
                "this ConvNormAct configuration (groups/dilation/norm='none'/non-ReLU activation) is outside the "
                "B200-native Darknet/VoVNet path and has no CUDA implementation here"
            )
        return super().forward(x)


# the north_star calls the unit "ConvBnAct"; the reference class is ConvNormAct
ConvBnAct = ConvNormAct
