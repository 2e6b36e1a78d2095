"""oracle_sinf/oracle_cosf (restated glibc 2.39 sinf/cosf) against the libm of the machine the test runs on.

The quick test samples; the `slow` test runs oracle/exhaustive_trig over EVERY binary32 value with |x| <= 4
(the reachable domain of CartPole's theta and MountainCar's 3*position) -- about 20 s on 8 cores."""
import ctypes
import ctypes.util
import os
import subprocess

import numpy as np
import pytest


def test_trig_matches_libm_sampled(oracle):
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.sinf.restype = libm.cosf.restype = ctypes.c_float
    libm.sinf.argtypes = libm.cosf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(11)
    xs = np.concatenate([
        rng.uniform(-0.3, 0.3, 20000), rng.uniform(-4, 4, 20000), rng.uniform(-130, 130, 20000),
        rng.standard_normal(10000) * 1e5, rng.standard_normal(5000) * 1e30,
        [0.0, -0.0, 2.0 ** -12, 0.78539816, 120.0, 3.4e38, 1e-40],
    ]).astype(np.float32)
    for x in xs:
        x = float(x)
        assert np.float32(oracle.sinf(x)).view(np.uint32) == np.float32(libm.sinf(x)).view(np.uint32), x
        assert np.float32(oracle.cosf(x)).view(np.uint32) == np.float32(libm.cosf(x)).view(np.uint32), x
    assert np.isnan(oracle.sinf(float("inf"))) and np.isnan(oracle.cosf(float("nan")))


@pytest.mark.slow
def test_trig_matches_libm_exhaustive(oracle):
    exe = os.path.join(os.path.dirname(oracle.LIB_PATH), "exhaustive_trig")
    out = subprocess.run([exe, "4", str(os.cpu_count() or 1)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
