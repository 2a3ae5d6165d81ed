"""Backbones of the hot path (reference vision_toolbox/backbones/__init__.py:3,10 exports Darknet,
DarknetYOLOv5, VoVNet; the transformer / torchvision wrappers of the reference are out of scope)."""
from .base import BaseBackbone
from .darknet import *  # noqa: F401,F403
from .darknet import Darknet, DarknetYOLOv5
from .vovnet import *  # noqa: F401,F403
from .vovnet import VoVNet
