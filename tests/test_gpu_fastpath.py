"""The branch-free fast forms equal the reference forms wherever their precondition holds.

Checked ON THE DEVICE, exhaustively where the domain is one float (all 2^32 bit patterns):
  * fdiv_const_fast(x, total_mass)  ==  __fdiv_rn(x, total_mass)   for every div_safe(x)
  * sincos_small(x)                 ==  sincos_ref(x)              for every |x| < 0.75
  * cos_fast(x), sincos_fast(x)     ==  cos_ref(x), sincos_ref(x)  for every |x| < 120
  * fmod_fast(x, 2 pi)              ==  fmodf(x, 2 pi)             for every |x| < 2^22
and on 2^29 Philox-drawn pairs for the two-operand fdiv_fast(a, b) == __fdiv_rn(a, b).
sincos_ref / cos_ref themselves are pinned to the oracle (and so to glibc) in test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import modurl_gym_b200 as m

    torch.cuda.init()
    return m.load_library()


@pytest.mark.parametrize("mode,name,min_checked", [
    (0, "fdiv_const_fast", 2 * 120 * (1 << 23)),
    (1, "sincos_small", 2 * 0x3f400000 - 10),
    (2, "cos_fast", 2 * 0x42f00000 - 10),
    (3, "fmod_fast", 2 * 0x4a800000 - 10),
    (4, "sincos_fast", 2 * 0x42f00000 - 10),
])
def test_fast_form_exhaustive(lib, mode, name, min_checked):
    out = (C.c_uint64 * 3)()
    checked = bad = 0
    first_bad = None
    for chunk in range(4):
        assert lib.mgym_probe_fast_exhaustive(mode, chunk << 30, 1 << 30, out) == 0
        checked += out[0]
        bad += out[1]
        if out[1] and first_bad is None:
            first_bad = out[2]
    assert bad == 0, f"{name}: {bad} of {checked} inputs differ, first bit pattern 0x{first_bad:08x}"
    assert checked >= min_checked, (name, checked)


def test_fdiv_fast_random_pairs(lib):
    out = (C.c_uint64 * 2)()
    assert lib.mgym_probe_fast_div_random(0xD1CE, 1 << 28, out) == 0
    assert out[0] > (1 << 28) and out[1] == 0, f"fdiv_fast: {out[1]} of {out[0]} pairs differ"


def test_cartpole_fast_forms_on_random_states(lib):
    """Env<0>::fast_ok (|theta| < 0.25, |theta_dot| < 10) is a SUFFICIENT precondition: on 2^28 random states drawn
    across its boundary the scalar and the packed-pair fast forms accept exactly the states inside it, and every
    accepted state steps to the reference form's bits (cartpole.rs:253-290 operator order)."""
    out = (C.c_uint64 * 3)()
    assert lib.mgym_probe_cartpole_fast(0xCA47, 1 << 27, out) == 0
    checked, bad, accepted = out[0], out[1], out[2]
    assert checked == 1 << 28 and bad == 0, f"CartPole fast forms: {bad} of {checked} states differ"
    assert 0.55 * checked < accepted < 0.85 * checked, accepted  # both sides of the precondition were exercised
