"""Checkpoint tooling: key renaming for ultralytics YOLOv5 backbones, release-file naming, strict loads."""
import pytest
import torch

from vision_toolbox_b200 import backbones, checkpoints as ck


def test_yolov5_key_rules_of_the_reference_script():
    # scripts/convert_yolov5_weights.py:10-16
    rules = {
        "stem.conv.weight": "model.0.conv.weight",
        "stages.0.conv.norm.bias": "model.1.norm.bias",
        "stages.0.conv1.conv.weight": "model.2.cv2.conv.weight",
        "stages.0.conv2.conv.weight": "model.2.cv1.conv.weight",
        "stages.0.blocks.1.conv2.norm.running_var": "model.2.m.1.cv2.norm.running_var",
        "stages.0.out_conv.norm.num_batches_tracked": "model.2.cv3.norm.num_batches_tracked",
        "stages.3.blocks.0.conv1.conv.weight": "model.8.m.0.cv1.conv.weight",
        "stages.2.conv.conv.weight": "model.5.conv.weight",
    }
    for ours, theirs in rules.items():
        assert ck.yolov5_key_to_ultralytics(ours) == theirs
        assert ck.yolov5_key_from_ultralytics(theirs) == ours
    assert ck.yolov5_key_to_ultralytics("stages.1.conv.norm.weight", rename_leaves=True) == "model.3.bn.weight"
    assert ck.yolov5_key_from_ultralytics("model.3.bn.weight", rename_leaves=True) == "stages.1.conv.norm.weight"
    for bad in ("head.weight", "stages.0.foo.conv.weight", "stages.x.conv.conv.weight"):
        with pytest.raises(ValueError):
            ck.yolov5_key_to_ultralytics(bad)
    with pytest.raises(ValueError):
        ck.yolov5_key_from_ultralytics("model.2.cv9.conv.weight")


@pytest.mark.parametrize("variant", ["n", "m"])
def test_yolov5_state_dict_round_trip(variant):
    m = backbones.DarknetYOLOv5.from_config(variant)
    sd = m.state_dict()
    there = ck.convert_yolov5_state_dict(sd, "ultralytics", rename_leaves=True)
    assert len(there) == len(sd) and all(k.startswith("model.") for k in there)
    assert {int(k.split(".")[1]) for k in there} == set(range(9))           # modules 0..8 of the ultralytics backbone
    back = ck.convert_yolov5_state_dict(there, "toolbox", rename_leaves=True)
    assert list(back) == list(sd) and all(back[k] is sd[k] for k in sd)
    m2 = backbones.DarknetYOLOv5.from_config(variant)
    m2.load_state_dict(back, strict=True)


def test_release_file_round_trip(tmp_path):
    torch.manual_seed(0)
    m = backbones.darknet19()
    classifier = torch.nn.Sequential(m, torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(1024, 10))
    lightning_like = {"model." + k: v for k, v in classifier.state_dict().items()}     # classifier.py:59-64
    sd = ck.backbone_state_dict_from_classifier(lightning_like)
    assert list(sd) == list(m.state_dict())
    path = ck.save_release_checkpoint(sd, "darknet19", str(tmp_path))
    assert path.endswith(".pth") and len(path.rsplit("-", 1)[1]) == 8 + 4
    m2 = backbones.darknet19()
    ck.load_backbone_checkpoint(m2, path)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    bad = path.replace(path.rsplit("-", 1)[1], "00000000.pth")
    import shutil

    shutil.copy(path, bad)
    with pytest.raises(ValueError, match="does not match the hash"):
        ck.load_backbone_checkpoint(m2, bad)
    with pytest.raises(ValueError):
        ck.backbone_state_dict_from_classifier({"foo": torch.zeros(1)})
