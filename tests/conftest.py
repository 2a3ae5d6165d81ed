import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 — the error metric of SURVEY.md §7 step 0."""
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


GOLDEN_CASES = sorted(p.stem for p in GOLDEN.glob("*.pt"))


def load_golden(name: str) -> dict:
    return torch.load(GOLDEN / f"{name}.pt", map_location="cpu", weights_only=False)


def case_kind(name: str) -> str:
    if name.startswith("model_vovnet"):
        return "vovnet"
    if name.startswith("model_yolov5"):
        return "yolov5"
    if name.startswith("model_"):
        return "darknet"
    return name.split("_")[0]  # unit / block / stage
