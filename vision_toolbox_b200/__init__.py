"""vision_toolbox_b200 — B200-native (sm_100a) drop-in for the ConvNormAct / Darknet / VoVNet path of
gau-nernst/vision-toolbox.  ``from vision_toolbox_b200 import backbones`` mirrors
``from vision_toolbox import backbones`` for that path; see INTEGRATION.md for aliasing the package name.
"""
from . import backbones, components
from .backbones import *  # noqa: F401,F403
from .components import *  # noqa: F401,F403
from .engine import get_precision, precision, set_precision  # noqa: F401  (bf16 tensor-core path | fp32 parity mode)

__version__ = "0.1.0"


def install_as(name: str = "vision_toolbox") -> None:
    """Register this package under another import name (e.g. the reference's) in ``sys.modules``."""
    import sys

    me = sys.modules[__name__]
    sys.modules[name] = me
    sys.modules[name + ".backbones"] = backbones
    sys.modules[name + ".components"] = components
    for sub in ("base", "darknet", "vovnet"):
        sys.modules[f"{name}.backbones.{sub}"] = sys.modules[f"{__name__}.backbones.{sub}"]
