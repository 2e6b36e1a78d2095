"""mgym_stats_allreduce with a raw ncclComm_t on two GPUs (tests/cpp/test_nccl_stats.cpp).  Skipped on one GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_nccl_stats_allreduce():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "test_nccl_stats"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_nccl_stats")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
