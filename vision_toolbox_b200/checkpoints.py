"""Checkpoint tooling around the drop-in state_dict layout - host-side only, no kernels.

What the reference ships for this path:
  * released backbone checkpoints named ``<variant>-<sha256[:8]>.pth`` holding the backbone's plain state_dict
    (``backbones/darknet.py:17``, ``vovnet.py:17``; produced by ``extras.py:112-128`` from a Lightning checkpoint whose
    classifier is ``nn.Sequential(backbone, pool, flatten, linear)`` -> keys prefixed ``model.0.``);
  * ``scripts/convert_yolov5_weights.py``: renaming between this layout and the ``model.<i>`` numbering of the
    ultralytics YOLOv5 backbone (stem -> model.0, stage k: conv -> model.{2k+1}, C3 block -> model.{2k+2} with
    conv1 <-> cv2, conv2 <-> cv1, blocks.<i>.conv<j> <-> m.<i>.cv<j>, out_conv <-> cv3).

Because the native modules keep the reference's parameter names, shapes and dtypes, none of this needs a kernel: a
checkpoint is loaded with ``load_state_dict(strict=True)`` and the bf16 operand packs are rebuilt on the next forward.
"""
from __future__ import annotations

import hashlib
import io
import os
import re
from typing import Mapping

import torch

__all__ = ["yolov5_key_to_ultralytics", "yolov5_key_from_ultralytics", "convert_yolov5_state_dict",
           "backbone_state_dict_from_classifier", "save_release_checkpoint", "load_backbone_checkpoint"]

# (toolbox module path inside a stage, ultralytics module path inside the C3 block)
_C3_PARTS = (("conv1", "cv2"), ("conv2", "cv1"), ("out_conv", "cv3"))
_LEAF_RENAMES = (("norm", "bn"),)   # ConvNormAct child names (components.py:26-39) vs ultralytics Conv


def _leaf(parts: list[str], to_ultralytics: bool, rename_leaves: bool) -> list[str]:
    if not rename_leaves:
        return parts
    table = dict(_LEAF_RENAMES) if to_ultralytics else {b: a for a, b in _LEAF_RENAMES}
    return [table.get(p, p) for p in parts]


def yolov5_key_to_ultralytics(key: str, rename_leaves: bool = False) -> str:
    """``stem.conv.weight`` -> ``model.0.conv.weight``; ``stages.1.blocks.2.conv1.norm.bias`` -> ``model.4.m.2.cv1.norm.bias``.

    `rename_leaves` additionally maps the ConvNormAct child ``norm`` to ultralytics' ``bn``."""
    parts = key.split(".")
    if parts[0] == "stem":
        return ".".join(["model", "0"] + _leaf(parts[1:], True, rename_leaves))
    if parts[0] != "stages" or len(parts) < 4 or not parts[1].isdigit():
        raise ValueError(f"not a DarknetYOLOv5 parameter name: {key}")
    stage, sub = int(parts[1]), parts[2]
    if sub == "conv":
        return ".".join(["model", str(2 * stage + 1)] + _leaf(parts[3:], True, rename_leaves))
    block = ["model", str(2 * stage + 2)]
    for ours, theirs in _C3_PARTS:
        if sub == ours:
            return ".".join(block + [theirs] + _leaf(parts[3:], True, rename_leaves))
    if sub == "blocks" and len(parts) >= 6 and re.fullmatch(r"conv[12]", parts[4]):
        return ".".join(block + ["m", parts[3], "cv" + parts[4][-1]] + _leaf(parts[5:], True, rename_leaves))
    raise ValueError(f"not a DarknetYOLOv5 parameter name: {key}")


def yolov5_key_from_ultralytics(key: str, rename_leaves: bool = False) -> str:
    """Inverse of :func:`yolov5_key_to_ultralytics` (backbone modules ``model.0`` ... ``model.8`` only)."""
    parts = key.split(".")
    if len(parts) < 3 or parts[0] != "model" or not parts[1].isdigit():
        raise ValueError(f"not an ultralytics YOLOv5 backbone parameter name: {key}")
    idx = int(parts[1])
    if idx == 0:
        return ".".join(["stem"] + _leaf(parts[2:], False, rename_leaves))
    stage = (idx - 1) // 2
    if idx % 2 == 1:
        return ".".join(["stages", str(stage), "conv"] + _leaf(parts[2:], False, rename_leaves))
    sub = parts[2]
    for ours, theirs in _C3_PARTS:
        if sub == theirs:
            return ".".join(["stages", str(stage), ours] + _leaf(parts[3:], False, rename_leaves))
    if sub == "m" and len(parts) >= 6 and re.fullmatch(r"cv[12]", parts[4]):
        return ".".join(["stages", str(stage), "blocks", parts[3], "conv" + parts[4][-1]] + _leaf(parts[5:], False, rename_leaves))
    raise ValueError(f"not an ultralytics YOLOv5 backbone parameter name: {key}")


def convert_yolov5_state_dict(sd: Mapping[str, torch.Tensor], to: str = "ultralytics",
                              rename_leaves: bool = False) -> dict[str, torch.Tensor]:
    """Rename every entry of a DarknetYOLOv5 state_dict (``to="ultralytics"``) or back (``to="toolbox"``)."""
    if to not in ("ultralytics", "toolbox"):
        raise KeyError(to)
    fn = yolov5_key_to_ultralytics if to == "ultralytics" else yolov5_key_from_ultralytics
    out = {fn(k, rename_leaves): v for k, v in sd.items()}
    if len(out) != len(sd):
        raise ValueError("key collision while renaming")
    return out


def backbone_state_dict_from_classifier(state_dict: Mapping[str, torch.Tensor], prefix: str = "model.0.") -> dict:
    """The backbone's entries of a classifier checkpoint (``classifier.py:59-64``: backbone is element 0 of the
    ``nn.Sequential`` stored as ``self.model``), prefix stripped - what ``extras.py:112-128`` releases."""
    out = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
    if not out:
        raise ValueError(f"no entries with prefix {prefix!r}")
    return out


def save_release_checkpoint(state_dict: Mapping[str, torch.Tensor], name: str, directory: str | None = None) -> str:
    """Write ``<name>-<first 8 hex digits of the file's sha256>.pth`` (the naming of the reference's releases)."""
    buf = io.BytesIO()
    torch.save({k: v.detach().cpu() for k, v in state_dict.items()}, buf)
    data = buf.getvalue()
    path = os.path.join(directory or os.getcwd(), f"{name}-{hashlib.sha256(data).hexdigest()[:8]}.pth")
    with open(path, "wb") as f:
        f.write(data)
    return path


def load_backbone_checkpoint(model: torch.nn.Module, path: str, check_hash: bool = True) -> None:
    """Strict load of a released checkpoint file; verifies the ``-<sha256[:8]>`` suffix of its name when present."""
    with open(path, "rb") as f:
        data = f.read()
    m = re.search(r"-([0-9a-f]{8})\.pth$", os.path.basename(path))
    if check_hash and m and hashlib.sha256(data).hexdigest()[:8] != m.group(1):
        raise ValueError(f"{path}: content does not match the hash in its name")
    model.load_state_dict(torch.load(io.BytesIO(data), map_location="cpu"), strict=True)
