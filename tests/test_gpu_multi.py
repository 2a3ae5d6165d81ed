"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on the single-GPU test box): launches tools/dp_check.py under torchrun.
SyncBN + gradient mean over 2 ranks == one process on the concatenated batch; peer-memory exchange == NCCL exchange."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_parity_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(ROOT / "tools" / "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FAIL" not in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_parity_against_the_oracle():
    """R ranks (all visible GPUs, at most 8) vs oracle/vt_oracle.py on the concatenated batch: tests/dp_parity.py."""
    n = min(8, torch.cuda.device_count())
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29543", str(ROOT / "tests" / "dp_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FAIL" not in out.stdout
