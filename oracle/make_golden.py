"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference) on CPU.  TEST INFRASTRUCTURE.

Run in the build container only (the reference tree does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For every case: a seeded reference module with randomised BN affine parameters and running statistics
(γ=1, β=0 would be a weak test), a seeded input, then
  * train-mode fp32 forward feature maps, gradients of  L = Σ_i <out_i, cot_i>  w.r.t. input and parameters,
    BN buffers after that step;
  * eval-mode fp32 forward;
  * train-mode forward + gradients under torch.autocast(bfloat16) (the "bf16 mode" oracle).
Also dumps the state_dict layout (keys, shapes, dtypes) of every released variant.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import torch

REF = os.environ.get("VT_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from vision_toolbox.backbones import Darknet, DarknetYOLOv5, VoVNet  # noqa: E402
from vision_toolbox.backbones.darknet import CSPDarknetStage, DarknetBlock, DarknetStage  # noqa: E402
from vision_toolbox.backbones.vovnet import OSABlock  # noqa: E402
from vision_toolbox.components import ConvNormAct  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def randomize_bn(m: torch.nn.Module, gen: torch.Generator) -> None:
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            with torch.no_grad():
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=gen) + 0.5)
                mod.bias.copy_(torch.rand(mod.bias.shape, generator=gen) * 0.4 - 0.2)
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) + 0.5)


def as_list(o):
    return list(o) if isinstance(o, (list, tuple)) else [o]


def run_case(name: str, build, x_shape, feature_fn=None, seed: int = 0):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    m = build()
    randomize_bn(m, gen)
    fwd = feature_fn or (lambda mod, x: mod(x))
    x = torch.rand(x_shape, generator=gen)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    rec = {"name": name, "x": x, "state_dict": sd0}

    # ---- fp32 train step
    m.train()
    xg = x.clone().requires_grad_(True)
    outs = as_list(fwd(m, xg))
    cots = [torch.randn(o.shape, generator=gen) for o in outs]
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    loss.backward()
    rec["cotangents"] = cots
    rec["train_fp32_outs"] = [o.detach().clone() for o in outs]
    rec["train_fp32_dx"] = xg.grad.clone()
    rec["train_fp32_dparams"] = {k: p.grad.clone() for k, p in m.named_parameters()}
    rec["buffers_after_step"] = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}

    # ---- fp32 eval forward (from the ORIGINAL buffers)
    m.load_state_dict(sd0)
    m.eval()
    with torch.no_grad():
        rec["eval_fp32_outs"] = [o.clone() for o in as_list(fwd(m, x))]

    # ---- bf16 autocast train step
    m.load_state_dict(sd0)
    m.train()
    m.zero_grad(set_to_none=True)
    xg = x.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        outs = as_list(fwd(m, xg))
        loss = sum((o.float() * c).sum() for o, c in zip(outs, cots))
    loss.backward()
    rec["train_bf16_outs"] = [o.detach().float().clone() for o in outs]
    rec["train_bf16_dx"] = xg.grad.clone()
    rec["train_bf16_dparams"] = {k: p.grad.clone() for k, p in m.named_parameters()}
    rec["buffers_after_bf16_step"] = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
    torch.save(rec, OUT / f"{name}.pt")
    n_par = sum(p.numel() for p in m.parameters())
    print(f"{name}: params {n_par}, outs {[tuple(o.shape) for o in outs]}, file {(OUT / (name + '.pt')).stat().st_size / 1e3:.0f} kB")


def main():
    fm = lambda mod, x: mod.get_feature_maps(x)
    # ConvNormAct units: every (k, s, p) class of SURVEY.md §8 a1 incl. the 3-channel stems and odd sizes
    run_case("unit_1x1_32_64", lambda: ConvNormAct(32, 64, 1), (2, 32, 9, 9))
    run_case("unit_3x3s1_16_32", lambda: ConvNormAct(16, 32), (2, 16, 11, 11))
    run_case("unit_3x3s2_32_32_odd", lambda: ConvNormAct(32, 32, 3, 2), (2, 32, 11, 11))
    run_case("unit_3x3s2_16_32_even", lambda: ConvNormAct(16, 32, 3, 2), (2, 16, 12, 12))
    run_case("unit_stem3x3_3_32", lambda: ConvNormAct(3, 32), (2, 3, 16, 16))
    run_case("unit_stem6x6s2_3_16", lambda: ConvNormAct(3, 16, 6, 2), (2, 3, 20, 20))
    # blocks
    run_case("block_darknet_32", lambda: DarknetBlock(32), (2, 32, 10, 10))
    run_case("stage_darknet_2_16_32", lambda: DarknetStage(2, 16, 32), (2, 16, 12, 12))
    run_case("stage_csp_2_16_32", lambda: CSPDarknetStage(2, 16, 32), (2, 16, 12, 12))
    # whole (narrow) models through the public factories' classes
    run_case("model_darknet", lambda: Darknet(16, [(0, 16), (1, 32), (2, 32)]), (2, 3, 32, 32), fm)
    run_case("model_cspdarknet", lambda: Darknet(16, [(1, 32), (2, 32)], CSPDarknetStage), (2, 3, 32, 32), fm)
    run_case("model_yolov5", lambda: DarknetYOLOv5(16, [(1, 32), (2, 32)]), (2, 3, 32, 32), fm)
    run_case("model_vovnet_ese", lambda: VoVNet(32, [(1, 16, 2, 32), (2, 16, 3, 32)], ese=True), (2, 3, 32, 32), fm)
    run_case("model_vovnet_v1", lambda: VoVNet(32, [(1, 16, 2, 48), (1, 16, 2, 48)], ese=False), (2, 3, 32, 32), fm)

    # state_dict layouts of the released variants
    layouts = {}
    variants = {f"darknet:{v}": (lambda v=v: Darknet.from_config(v)) for v in ("darknet19", "darknet53", "cspdarknet53")}
    variants.update({f"yolov5:{v}": (lambda v=v: DarknetYOLOv5.from_config(v)) for v in "nsmlx"})
    for v, slim, e in [(27, True, False), (39, False, False), (57, False, False), (19, True, True), (19, False, True),
                       (39, False, True), (57, False, True), (99, False, True)]:
        variants[f"vovnet:{v}:{int(slim)}:{int(e)}"] = (lambda v=v, slim=slim, e=e: VoVNet.from_config(v, slim, e))
    for name, build in variants.items():
        with torch.device("meta"):
            m = build()
        layouts[name] = {
            "keys": [[k, list(t.shape), str(t.dtype)] for k, t in m.state_dict().items()],
            "out_channels_list": list(m.out_channels_list),
            "stride": m.stride,
            "n_params": sum(p.numel() for p in m.parameters()),
        }
    (OUT / "state_dict_layouts.json").write_text(json.dumps(layouts))
    print("layouts:", {k: len(v["keys"]) for k, v in layouts.items()})


if __name__ == "__main__":
    main()
