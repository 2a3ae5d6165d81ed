"""BaseBackbone — drop-in for reference vision_toolbox/backbones/base.py:14-25."""
from __future__ import annotations

from abc import ABCMeta, abstractmethod

import torch
from torch import Tensor, nn


class BaseBackbone(nn.Module, metaclass=ABCMeta):
    """Subclasses describe their dataflow twice, from the same modules and parameters:

    * ``_features_cpu(x)``: plain torch composition for CPU tensors (shape tests, jit.trace);
    * ``_emit(graph, x)``: the same dataflow emitted into the native planner for CUDA tensors.
    """

    out_channels_list: tuple[int, ...]
    stride: int

    @abstractmethod
    def _features_cpu(self, x: Tensor) -> list[Tensor]:
        pass

    @abstractmethod
    def _emit(self, g, x):
        pass

    def get_feature_maps(self, x: Tensor) -> list[Tensor]:
        if x.is_cuda:
            from ..engine import run_native

            return run_native(self, x)
        return self._features_cpu(x)

    def forward(self, x: Tensor) -> Tensor:
        return self.get_feature_maps(x)[-1]

    def get_last_out_channels(self) -> int:
        # used by the reference trainer (classifier.py:63)
        return self.out_channels_list[-1]

    def _load_state_dict_from_url(self, url: str) -> None:
        state_dict = torch.hub.load_state_dict_from_url(url)
        self.load_state_dict(state_dict)
