"""Builders shared by the CPU and GPU tests: golden-case name -> our module / oracle call."""
from __future__ import annotations

import torch

from oracle import vt_oracle as O
from vision_toolbox_b200.backbones import Darknet, DarknetYOLOv5, VoVNet
from vision_toolbox_b200.backbones.darknet import CSPDarknetStage, DarknetBlock, DarknetStage
from vision_toolbox_b200.components import ConvNormAct

# the same constructor calls oracle/make_golden.py made on the reference classes
BUILDERS = {
    "unit_1x1_32_64": lambda: ConvNormAct(32, 64, 1),
    "unit_3x3s1_16_32": lambda: ConvNormAct(16, 32),
    "unit_3x3s2_32_32_odd": lambda: ConvNormAct(32, 32, 3, 2),
    "unit_3x3s2_16_32_even": lambda: ConvNormAct(16, 32, 3, 2),
    "unit_stem3x3_3_32": lambda: ConvNormAct(3, 32),
    "unit_stem6x6s2_3_16": lambda: ConvNormAct(3, 16, 6, 2),
    "block_darknet_32": lambda: DarknetBlock(32),
    "stage_darknet_2_16_32": lambda: DarknetStage(2, 16, 32),
    "stage_csp_2_16_32": lambda: CSPDarknetStage(2, 16, 32),
    "model_darknet": lambda: Darknet(16, [(0, 16), (1, 32), (2, 32)]),
    "model_cspdarknet": lambda: Darknet(16, [(1, 32), (2, 32)], CSPDarknetStage),
    "model_yolov5": lambda: DarknetYOLOv5(16, [(1, 32), (2, 32)]),
    "model_vovnet_ese": lambda: VoVNet(32, [(1, 16, 2, 32), (2, 16, 3, 32)], ese=True),
    "model_vovnet_v1": lambda: VoVNet(32, [(1, 16, 2, 48), (1, 16, 2, 48)], ese=False),
}
UNIT_STRIDE = {"unit_3x3s2_32_32_odd": 2, "unit_3x3s2_16_32_even": 2, "unit_stem6x6s2_3_16": 2}


def module_outputs(m, x):
    from vision_toolbox_b200.backbones.base import BaseBackbone

    out = m.get_feature_maps(x) if isinstance(m, BaseBackbone) else m(x)
    return list(out) if isinstance(out, (list, tuple)) else [out]


def oracle_outputs(name: str, sd: dict, x: torch.Tensor, training: bool, mode: str, new_stats=None):
    """Run the functional oracle for a golden case."""
    sd = dict(sd)
    if name.startswith("unit_"):
        sd["__stride__"] = {"": UNIT_STRIDE.get(name, 1)}
        return [O.conv_norm_act(sd, "", x, training, mode, new_stats=new_stats)]
    if name.startswith("block_darknet"):
        return [O.darknet_block(sd, "", x, training, mode, new_stats)]
    if name.startswith("stage_darknet"):
        sd["__stride__"] = {"conv.": 2}
        return [O.darknet_stage(sd, "", x, training, mode, new_stats)]
    if name.startswith("stage_csp"):
        sd["__stride__"] = {"conv.": 2}
        return [O.csp_stage(sd, "", x, training, mode, new_stats)]
    kind = "vovnet" if "vovnet" in name else ("yolov5" if "yolov5" in name else "darknet")
    return O.features(kind, sd, x, training, mode, new_stats)


def oracle_step(name, sd, x, cots, mode):
    """Train-mode forward + backward of L = sum <out_i, cot_i> through the oracle (autograd over its formulas)."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    xg = x.clone().requires_grad_(True)
    new_stats = {}
    outs = oracle_outputs(name, full, xg, True, mode, new_stats)
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    grads = torch.autograd.grad(loss, [xg] + list(params.values()), allow_unused=True)
    return outs, grads[0], dict(zip(params.keys(), grads[1:])), new_stats
