"""The oracle (oracle/vt_oracle.py) pinned against vectors produced by the unmodified reference
(oracle/make_golden.py ran /root/reference on CPU).  CPU only."""
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden, rel_err
from helpers import oracle_outputs, oracle_step

FP32_FWD_TOL = 2e-6   # same torch CPU kernels, different op order in BN only
FP32_GRAD_TOL = 5e-5
BF16_FWD_TOL = 2e-2   # north_star tolerance for bf16 mode
BF16_STATS_TOL = 2e-3


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_fp32_train_forward_backward(name):
    g = load_golden(name)
    outs, dx, dparams, new_stats = oracle_step(name, g["state_dict"], g["x"], g["cotangents"], "fp32")
    for o, ref in zip(outs, g["train_fp32_outs"]):
        assert o.shape == ref.shape
        assert rel_err(o, ref) < FP32_FWD_TOL
    assert rel_err(dx, g["train_fp32_dx"]) < FP32_GRAD_TOL
    for k, ref in g["train_fp32_dparams"].items():
        assert rel_err(dparams[k], ref) < FP32_GRAD_TOL, k
    for k, ref in g["buffers_after_step"].items():
        if "num_batches" in k:
            assert int(new_stats[k]) == int(ref)
        else:
            assert rel_err(new_stats[k], ref) < 1e-5, k   # north_star: BN running statistics within 1e-5


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_fp32_eval_forward(name):
    g = load_golden(name)
    with torch.no_grad():
        outs = oracle_outputs(name, g["state_dict"], g["x"], False, "fp32")
    for o, ref in zip(outs, g["eval_fp32_outs"]):
        assert rel_err(o, ref) < FP32_FWD_TOL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_bf16_mode_matches_reference_autocast(name):
    """bf16 mode restates autocast's rounding points; the reference's own autocast run is the yardstick."""
    g = load_golden(name)
    new_stats = {}
    with torch.no_grad():
        outs = oracle_outputs(name, g["state_dict"], g["x"], True, "bf16", new_stats)
    for o, ref in zip(outs, g["train_bf16_outs"]):
        assert rel_err(o, ref) < BF16_FWD_TOL
    for k, ref in g["buffers_after_bf16_step"].items():
        assert rel_err(new_stats[k], ref) < BF16_STATS_TOL, k
