"""CPU oracle for the ConvNormAct / Darknet / VoVNet hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this module; the product package (vision_toolbox_b200) never does.

It restates, as *stateless functions over a state_dict*, what the reference computes:

* ``conv_norm_act``   — vision_toolbox/components.py:13-46  (Conv2d(bias=False) → BatchNorm2d → ReLU)
* ``darknet_block``   — backbones/darknet.py:20-28           (x + conv2(conv1(x)))
* ``darknet_stage``   — backbones/darknet.py:31-36
* ``csp_stage``       — backbones/darknet.py:39-55           (cat([conv1(o), blocks(conv2(o))]) → out_conv)
* ``darknet_features`` / ``yolov5_features`` — darknet.py:83-87 / :116-120
* ``ese`` / ``osa_block`` / ``vovnet_features`` — backbones/vovnet.py:20-28 / :31-63 / :100-104

The arithmetic itself lives in a third-party dependency the reference does not vendor or pin (torch ATen:
setup.cfg:10-12 lists plain ``torch``); the oracle of record is torch 2.11.0 CPU in this image.  BatchNorm is
written out explicitly (mean, biased variance, unbiased running update, eps inside the sqrt) rather than
calling ``F.batch_norm`` so the formulas the CUDA kernels implement are visible here.

PINNING: the reference's own tests hold no value fixtures for this path (tests/test_backbones.py:39-78 check
shapes only — "parity unpinned" by the reference itself).  This oracle is therefore pinned against outputs of
the reference run in the build container: ``oracle/make_golden.py`` imports /root/reference, and
``tests/test_oracle_golden.py`` checks every function here against those committed vectors (tests/golden/).

Two numeric modes:
* ``mode="fp32"``  — plain fp32, what ``model(x)`` does on CPU in the reference.
* ``mode="bf16"``  — the rounding points of ``torch.autocast(dtype=torch.bfloat16)`` around the reference
  (conv operands and result in bf16, BN statistics in fp32 from the bf16 tensor, BN result bf16, adds in bf16).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor

EPS = 1e-5
MOMENTUM = 0.1


def _r(t: Tensor, mode: str) -> Tensor:
    """Round to the activation dtype of the mode (differentiable: straight bf16 round-trip)."""
    if mode == "bf16":
        return t.to(torch.bfloat16).to(t.dtype)
    return t


def conv_norm_act(sd: dict, prefix: str, x: Tensor, training: bool, mode: str = "fp32", act: bool = True,
                  new_stats: dict | None = None) -> Tensor:
    """components.py:13-46.  `sd[prefix + 'conv.weight']`, `sd[prefix + 'norm.*']`."""
    w = sd[prefix + "conv.weight"]
    k = w.shape[-1]
    # stride is not stored in the state_dict: the reference fixes it by position (stage convs and the YOLOv5 /
    # VoVNet stems are stride 2); callers pass it through sd["__stride__"][prefix]
    stride = sd.get("__stride__", {}).get(prefix, 1)
    pad = math.ceil((k - stride) / 2)  # components.py:31
    y = F.conv2d(_r(x, mode), _r(w, mode), None, stride, pad)  # components.py:26-35, bias-free
    y = _r(y, mode)
    gamma, beta = sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]
    if training:
        m = y.numel() // y.shape[1]
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if new_stats is not None:
            with torch.no_grad():
                rm, rv = sd[prefix + "norm.running_mean"], sd[prefix + "norm.running_var"]
                new_stats[prefix + "norm.running_mean"] = (1 - MOMENTUM) * rm + MOMENTUM * mean
                new_stats[prefix + "norm.running_var"] = (1 - MOMENTUM) * rv + MOMENTUM * var * (m / max(m - 1, 1))
                new_stats[prefix + "norm.num_batches_tracked"] = sd[prefix + "norm.num_batches_tracked"] + 1
    else:
        mean, var = sd[prefix + "norm.running_mean"], sd[prefix + "norm.running_var"]
    invstd = torch.rsqrt(var + EPS)
    z = (y - mean[None, :, None, None]) * invstd[None, :, None, None] * gamma[None, :, None, None] + beta[None, :, None, None]
    z = _r(z, mode)
    return torch.relu(z) if act else z  # components.py:39, in-place ReLU


def darknet_block(sd, prefix, x, training, mode="fp32", new_stats=None):
    """darknet.py:27-28 — residual added AFTER conv2's ReLU."""
    h = conv_norm_act(sd, prefix + "conv1.", x, training, mode, new_stats=new_stats)
    h = conv_norm_act(sd, prefix + "conv2.", h, training, mode, new_stats=new_stats)
    return _r(x + h, mode)


def _n_blocks(sd, prefix):
    n = 0
    while f"{prefix}blocks.{n}.conv1.conv.weight" in sd:
        n += 1
    return n


def darknet_stage(sd, prefix, x, training, mode="fp32", new_stats=None):
    """darknet.py:31-36 (a stage with n_blocks == 0 is a bare stride-2 ConvNormAct, darknet.py:79)."""
    if prefix + "conv.conv.weight" not in sd:
        return conv_norm_act(sd, prefix, x, training, mode, new_stats=new_stats)
    x = conv_norm_act(sd, prefix + "conv.", x, training, mode, new_stats=new_stats)
    for j in range(_n_blocks(sd, prefix)):
        x = darknet_block(sd, f"{prefix}blocks.{j}.", x, training, mode, new_stats)
    return x


def csp_stage(sd, prefix, x, training, mode="fp32", new_stats=None):
    """darknet.py:51-55."""
    out = conv_norm_act(sd, prefix + "conv.", x, training, mode, new_stats=new_stats)
    a = conv_norm_act(sd, prefix + "conv1.", out, training, mode, new_stats=new_stats)
    b = conv_norm_act(sd, prefix + "conv2.", out, training, mode, new_stats=new_stats)
    for j in range(_n_blocks(sd, prefix)):
        b = darknet_block(sd, f"{prefix}blocks.{j}.", b, training, mode, new_stats)
    out = torch.cat([a, b], dim=1)
    return conv_norm_act(sd, prefix + "out_conv.", out, training, mode, new_stats=new_stats)


def _stage_any(sd, prefix, x, training, mode, new_stats):
    if prefix + "out_conv.conv.weight" in sd:
        return csp_stage(sd, prefix, x, training, mode, new_stats)
    return darknet_stage(sd, prefix, x, training, mode, new_stats)


def _n_stages(sd):
    n = 0
    while any(k.startswith(f"stages.{n}.") for k in sd):
        n += 1
    return n


def darknet_features(sd, x, training, mode="fp32", new_stats=None):
    """darknet.py:83-87 — the stem output is dropped."""
    outs = [conv_norm_act(sd, "stem.", x, training, mode, new_stats=new_stats)]
    for i in range(_n_stages(sd)):
        outs.append(_stage_any(sd, f"stages.{i}.", outs[-1], training, mode, new_stats))
    return outs[1:]


def yolov5_features(sd, x, training, mode="fp32", new_stats=None):
    """darknet.py:116-120 — stem (6x6 stride 2) output is kept."""
    outs = [conv_norm_act(sd, "stem.", x, training, mode, new_stats=new_stats)]
    for i in range(_n_stages(sd)):
        outs.append(csp_stage(sd, f"stages.{i}.", outs[-1], training, mode, new_stats))
    return outs


def ese(sd, prefix, x, mode="fp32"):
    """vovnet.py:20-28 — x * hardsigmoid(conv1x1_bias(avgpool(x)))."""
    pooled = _r(x.mean(dim=(2, 3), keepdim=True), mode)
    z = F.conv2d(pooled, _r(sd[prefix + "linear.weight"], mode), _r(sd[prefix + "linear.bias"], mode))
    z = _r(z, mode)
    gate = _r(torch.clamp(z / 6 + 0.5, 0, 1), mode)  # nn.Hardsigmoid
    return _r(x * gate, mode)


def osa_block(sd, prefix, x, training, mode="fp32", new_stats=None):
    """vovnet.py:50-63."""
    outs = [x]
    i = 0
    while f"{prefix}convs.{i}.conv.weight" in sd:
        outs.append(conv_norm_act(sd, f"{prefix}convs.{i}.", outs[-1], training, mode, new_stats=new_stats))
        i += 1
    out = conv_norm_act(sd, prefix + "out_conv.", torch.cat(outs, dim=1), training, mode, new_stats=new_stats)
    if prefix + "ese.linear.weight" in sd:
        out = ese(sd, prefix + "ese.", out, mode)
    if out.shape[1] == x.shape[1]:  # vovnet.py:48 residual iff in_channels == out_channels
        out = _r(out + x, mode)
    return out


def vovnet_features(sd, x, training, mode="fp32", new_stats=None):
    """vovnet.py:84-104 — 3-conv stem, then per stage MaxPool2d(3,2,1) + OSA blocks."""
    for i in range(3):
        x = conv_norm_act(sd, f"stem.{i}.", x, training, mode, new_stats=new_stats)
    outs = [x]
    for s in range(_n_stages(sd)):
        h = F.max_pool2d(outs[-1], 3, 2, 1)
        j = 0
        while any(k.startswith(f"stages.{s}.module_{j}.") for k in sd):
            h = osa_block(sd, f"stages.{s}.module_{j}.", h, training, mode, new_stats)
            j += 1
        outs.append(h)
    return outs


def stride_table(model_kind: str, sd: dict) -> dict:
    """Which ConvNormAct prefixes are stride 2 (fixed by position in the reference, not stored in weights)."""
    tbl = {}
    if model_kind in ("darknet", "yolov5"):
        if model_kind == "yolov5":
            tbl["stem."] = 2  # darknet.py:109
        for i in range(_n_stages(sd)):
            p = f"stages.{i}."
            tbl[p + "conv." if p + "conv.conv.weight" in sd else p] = 2  # darknet.py:34,43,79
    elif model_kind == "vovnet":
        tbl["stem.0."] = 2  # vovnet.py:85
    return tbl


def features(model_kind: str, sd: dict, x: Tensor, training: bool, mode: str = "fp32", new_stats=None):
    sd = dict(sd)
    sd["__stride__"] = stride_table(model_kind, sd)
    fn = dict(darknet=darknet_features, yolov5=yolov5_features, vovnet=vovnet_features)[model_kind]
    return fn(sd, x, training, mode, new_stats)


# ---------------------------------------------------------------------------------------------------
# training step of the reference trainer (classifier.py:59-64, 83-95): backbone → avg-pool → linear → CE
# ---------------------------------------------------------------------------------------------------
def classifier_loss(model_kind, sd, head_w, head_b, x, labels, mode="fp32", label_smoothing=0.1, new_stats=None):
    f = features(model_kind, sd, x, True, mode, new_stats)[-1]
    pooled = f.mean(dim=(2, 3))
    logits = F.linear(_r(pooled, mode), _r(head_w, mode), _r(head_b, mode))
    return F.cross_entropy(logits.float(), labels, label_smoothing=label_smoothing)
