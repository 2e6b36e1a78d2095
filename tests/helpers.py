"""Shared helpers for the parity tests (GPU path vs CPU oracle)."""
import numpy as np

KIND_NAMES = ["CartPole-v1", "MountainCar-v0", "MountainCarContinuous-v0", "Pendulum-v1", "Acrobot-v1"]
STATE_DIM = [4, 2, 2, 2, 4]
OBS_DIM = [4, 2, 2, 3, 6]
CONTINUOUS = [False, False, True, True, False]
NUM_ACTIONS = [2, 3, 0, 0, 3]


def bits(a):
    """Bit pattern view, so that comparisons are exact (distinguish -0.0, compare NaNs)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a


def assert_bit_equal(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    neq = bits(got) != bits(want)
    if neq.any():
        idx = np.argwhere(neq)[0]
        raise AssertionError(
            f"{what}: {int(neq.sum())} of {neq.size} elements differ; first at {tuple(idx)}: "
            f"got {got[tuple(idx)]!r} want {want[tuple(idx)]!r}")


def random_actions(rng, kind, shape):
    if CONTINUOUS[kind]:
        lim = 1.2 if kind == 2 else 2.5   # beyond the Box bounds on purpose: exercises the clamps
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)
    return rng.integers(0, NUM_ACTIONS[kind], size=shape, dtype=np.uint8)


def random_states(rng, kind, n):
    """Broad initial states (wider than reset()) so that clamps, walls and thresholds are hit."""
    sd = STATE_DIM[kind]
    s = np.zeros((sd, n), dtype=np.float32)
    if kind == 0:
        s[0] = rng.uniform(-2.5, 2.5, n)
        s[1] = rng.uniform(-3, 3, n)
        s[2] = rng.uniform(-0.25, 0.25, n)
        s[3] = rng.uniform(-3, 3, n)
    elif kind in (1, 2):
        s[0] = rng.uniform(-1.2, 0.6, n)
        s[1] = rng.uniform(-0.07, 0.07, n)
        s[0, : n // 16] = -1.2           # on the left wall
        s[0, n // 16: n // 8] = 0.49     # next to the goal
        s[1, n // 16: n // 8] = 0.05
    elif kind == 3:
        s[0] = rng.uniform(-10, 10, n)
        s[1] = rng.uniform(-8, 8, n)
    else:
        s[0] = rng.uniform(-np.pi, np.pi, n)
        s[1] = rng.uniform(-np.pi, np.pi, n)
        s[2] = rng.uniform(-12, 12, n)
        s[3] = rng.uniform(-28, 28, n)
    return s
