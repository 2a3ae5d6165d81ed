"""FPN / PAN (reference necks.py:45-120): CPU path bit-identical to the reference's outputs and gradients
(tests/golden_extras/necks.pt from oracle/make_golden_necks.py); CUDA path (bias unit + resize-fuse kernel + native
ConvNormAct) within the bf16 budget of the same vectors, also fed straight from a native backbone."""
from pathlib import Path

import pytest
import torch

from conftest import rel_err
from vision_toolbox_b200 import necks as N

G = torch.load(Path(__file__).resolve().parent / "golden_extras" / "necks.pt", map_location="cpu", weights_only=False)


def _build(c):
    m = getattr(N, c["cls"])([16, 32, 48], 16, **c["kw"]).train()
    assert list(m.state_dict().keys()) == list(c["state_dict"].keys())       # the reference's state_dict layout
    m.load_state_dict(c["state_dict"])
    return m


@pytest.mark.parametrize("name", list(G))
def test_cpu_path_is_the_reference(name):
    c = G[name]
    m = _build(c)
    xg = [x.clone().requires_grad_(True) for x in c["xs"]]
    outs = m(list(xg))
    sum((o * k).sum() for o, k in zip(outs, c["cots"])).backward()
    for o, r in zip(outs, c["outs"]):
        assert torch.equal(o, r)
    for x, r in zip(xg, c["dxs"]):
        assert torch.allclose(x.grad, r, rtol=1e-5, atol=1e-6)
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, c["dparams"][k], rtol=1e-4, atol=1e-5), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G))
def test_cuda_path_matches_the_reference_vectors(name):
    from vision_toolbox_b200 import _lib

    c = G[name]
    m = _build(c).cuda()
    xg = [x.cuda().requires_grad_(True) for x in c["xs"]]
    n0 = _lib.launch_count()
    outs = m(list(xg))
    assert _lib.launch_count() - n0 >= 8                      # the native kernels ran (no torch fallback for CUDA tensors)
    sum((o.float() * k.cuda()).sum() for o, k in zip(outs, c["cots"])).backward()
    torch.cuda.synchronize()
    for o, r in zip(outs, c["outs"]):
        assert tuple(o.shape) == tuple(r.shape)      # (an Identity lateral passes its input through, as in the reference)
        assert rel_err(o.float(), r) < 2e-2
    # gradients: small train-mode BatchNorm cases in bf16 against fp32 truth - the loose Appendix-B style bound
    for x, r in zip(xg, c["dxs"]):
        assert rel_err(x.grad, r) < 1.5e-1
    for k, p in m.named_parameters():
        assert torch.isfinite(p.grad).all() and rel_err(p.grad, c["dparams"][k]) < 2e-1, (k, rel_err(p.grad, c["dparams"][k]))


@pytest.mark.gpu
@pytest.mark.parametrize("k,stride", [(1, 1), (3, 1), (3, 2)])
def test_bias_conv_unit_on_gpu(k, stride):
    """A bare nn.Conv2d with bias through the planner's bias unit (bias in the conv epilogue; bias gradient = per-channel
    sum of the output gradient) against torch on the CPU, teacher-forced: the bf16 budget."""
    from vision_toolbox_b200.necks import _BiasConvUnit

    torch.manual_seed(0)
    conv = torch.nn.Conv2d(32, 48, k, stride, padding=k // 2)
    x = torch.rand(4, 32, 18, 20)
    xr = x.clone().requires_grad_(True)
    ref = conv(xr)
    cot = torch.randn_like(ref)
    (ref * cot).sum().backward()
    want = (ref.detach(), xr.grad, conv.weight.grad.clone(), conv.bias.grad.clone())
    conv.zero_grad()
    unit = _BiasConvUnit(conv).cuda()
    xg = x.cuda().requires_grad_(True)
    out = unit(xg)
    (out.float() * cot.cuda()).sum().backward()
    torch.cuda.synchronize()
    got = (out.float(), xg.grad, conv.weight.grad, conv.bias.grad)
    for name, a, b in zip(("out", "dx", "dW", "db"), got, want):
        assert rel_err(a, b) < 1e-2, (name, rel_err(a, b))


@pytest.mark.gpu
def test_neck_on_native_backbone_features():
    """FPN on the NHWC bf16 maps of a native backbone (the consumer of get_feature_maps): no layout
    conversion of the inputs, gradients flow back into the backbone."""
    from vision_toolbox_b200.backbones import Darknet
    from vision_toolbox_b200.backbones.darknet import CSPDarknetStage

    torch.manual_seed(0)
    bb = Darknet(16, [(1, 32), (1, 64), (1, 128)], CSPDarknetStage).cuda().train()
    neck = N.PAN(list(bb.out_channels_list), 64).cuda().train()
    x = torch.rand(2, 3, 64, 64, device="cuda")
    feats = bb.get_feature_maps(x)
    outs = neck(feats)
    assert [tuple(o.shape) for o in outs] == [(2, 64, 32, 32), (2, 64, 16, 16), (2, 64, 8, 8)]
    sum(o.float().square().mean() for o in outs).backward()
    torch.cuda.synchronize()
    for k, p in list(bb.named_parameters()) + list(neck.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    assert float(bb.stem.conv.weight.grad.abs().sum()) > 0
    # eval mode: fused epilogues, deterministic
    bb.eval(); neck.eval()
    with torch.no_grad():
        a = neck(bb.get_feature_maps(x))
        b = neck(bb.get_feature_maps(x))
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    # odd pyramid levels cannot be fused (the reference's `+` raises too)
    with pytest.raises(RuntimeError):
        N.FPN([16, 16], 16).cuda()([torch.rand(1, 16, 11, 11, device="cuda"), torch.rand(1, 16, 6, 6, device="cuda")])
