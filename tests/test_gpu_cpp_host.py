"""Runs the C++ host-layer parity test (tests/cpp/test_gpu_vec_env.cpp over include/mgym.hpp): the reference's
known-answer replay, its behavioural unit tests and a batched bit-equality check against the oracle, all from
C++ through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_layer_parity():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = os.path.join(ROOT, "tests", "cpp", "test_gpu_vec_env")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
    assert out.stdout.count("ok ") >= 10
