"""Drop-in surface of the reference, CPU only.

* state_dict layout (keys, shapes, dtypes), out_channels_list, stride and parameter counts of every released
  variant against tests/golden/state_dict_layouts.json (written by oracle/make_golden.py from /root/reference);
* the reference's own API tests (tests/test_backbones.py:39-78: attributes, forward, get_feature_maps, jit.trace)
  restated on our classes;
* the CPU composition of our modules against the reference's golden outputs (same nn modules -> same numbers)."""
import json

import pytest
import torch

from conftest import GOLDEN, GOLDEN_CASES, load_golden, rel_err
from helpers import BUILDERS, module_outputs
from vision_toolbox_b200 import backbones
from vision_toolbox_b200.backbones import Darknet, DarknetYOLOv5, VoVNet
from vision_toolbox_b200.components import ConvBnAct, ConvNormAct

LAYOUTS = json.loads((GOLDEN / "state_dict_layouts.json").read_text())


def _build(key: str):
    fam, *rest = key.split(":")
    if fam == "darknet":
        return Darknet.from_config(rest[0])
    if fam == "yolov5":
        return DarknetYOLOv5.from_config(rest[0])
    return VoVNet.from_config(int(rest[0]), bool(int(rest[1])), bool(int(rest[2])))


@pytest.mark.parametrize("key", sorted(LAYOUTS))
def test_state_dict_layout_matches_reference(key):
    ref = LAYOUTS[key]
    with torch.device("meta"):
        m = _build(key)
    ours = [[k, list(t.shape), str(t.dtype)] for k, t in m.state_dict().items()]
    assert ours == ref["keys"]
    assert list(m.out_channels_list) == ref["out_channels_list"] and isinstance(m.out_channels_list, tuple)
    assert m.stride == ref["stride"] and isinstance(m.stride, int)
    assert sum(p.numel() for p in m.parameters()) == ref["n_params"]
    assert m.get_last_out_channels() == ref["out_channels_list"][-1]


FACTORIES = ["darknet19", "darknet53", "cspdarknet53", "darknet_yolov5n", "darknet_yolov5s", "darknet_yolov5m",
             "darknet_yolov5l", "darknet_yolov5x", "vovnet27_slim", "vovnet39", "vovnet57", "vovnet19_slim_ese",
             "vovnet19_ese", "vovnet39_ese", "vovnet57_ese", "vovnet99_ese"]


def test_factory_functions_exist():
    for f in FACTORIES:
        assert callable(getattr(backbones, f)), f
    assert ConvBnAct is ConvNormAct


# reference tests/test_backbones.py:24-36 factories that are on the hot path
REF_TEST_FACTORIES = [
    lambda: Darknet.from_config("darknet19"),
    lambda: DarknetYOLOv5.from_config("n"),
    lambda: VoVNet.from_config(27, True),
    lambda: VoVNet.from_config(19, True, True),
]


@pytest.mark.parametrize("factory", REF_TEST_FACTORIES)
def test_reference_api_contract(factory):
    m = factory()
    x = torch.rand(1, 3, 64, 64)
    assert isinstance(m.out_channels_list, tuple) and all(isinstance(c, int) for c in m.out_channels_list)
    assert isinstance(m.stride, int) and callable(m.get_feature_maps)
    out = m(x)
    assert isinstance(out, torch.Tensor) and out.dim() == 4
    fms = m.get_feature_maps(x)
    assert isinstance(fms, list) and len(fms) == len(m.out_channels_list)
    for f, c in zip(fms, m.out_channels_list):
        assert f.shape[1] == c
    torch.jit.trace(m, x)   # tests/test_backbones.py:76-78 (train mode, CPU)


def test_error_behaviour():
    with pytest.raises(AssertionError):
        backbones.CSPDarknetStage(0, 16, 32)          # darknet.py:41
    with pytest.raises(AssertionError):
        Darknet(16, [])                               # darknet.py:70
    with pytest.raises(KeyError):
        Darknet.from_config("darknet99")
    with pytest.raises(KeyError):
        ConvNormAct(8, 8, act="tanh")
    with pytest.raises(KeyError):
        ConvNormAct(8, 8, norm="ln")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cpu_composition_matches_reference_golden(name):
    """CPU tensors take the nn.Module composition; it must reproduce the reference's fp32 results exactly-ish."""
    g = load_golden(name)
    m = BUILDERS[name]()
    m.load_state_dict(g["state_dict"])          # strict: same keys as the reference module
    m.train()
    outs = module_outputs(m, g["x"])
    for o, ref in zip(outs, g["train_fp32_outs"]):
        assert rel_err(o, ref) < 1e-6
    sd = m.state_dict()
    for k, ref in g["buffers_after_step"].items():
        assert rel_err(sd[k].float(), ref.float()) < 1e-6, k


def test_cuda_tensor_without_kernel_support_raises():
    m = ConvNormAct(16, 16, groups=2)
    assert not m.native_supported()
    m2 = ConvNormAct(16, 32)
    assert m2.native_supported()


def test_install_as_reference_name():
    import sys

    import vision_toolbox_b200

    vision_toolbox_b200.install_as("vision_toolbox_alias_for_test")
    from vision_toolbox_alias_for_test.backbones import Darknet as D2  # noqa: E402
    from vision_toolbox_alias_for_test.components import ConvNormAct as C2  # noqa: E402

    assert D2 is Darknet and C2 is ConvNormAct
    for k in [k for k in sys.modules if k.startswith("vision_toolbox_alias_for_test")]:
        del sys.modules[k]
