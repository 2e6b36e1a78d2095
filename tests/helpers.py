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


# ---- 10^4-step teacher-forced traces of the kinds the reference lacks (tests/golden/make_f64_traces.py) ----------
F64_TRACES = {0: "cartpole", 1: "mountain_car", 2: "mountain_car_continuous", 3: "pendulum", 4: "acrobot"}
# north_star: "within 1e-6 relative for CartPole/MountainCar ... within 1e-5 for Acrobot/Pendulum" -- relative, with
# an absolute floor of the same size
F64_TOL = {0: 1e-6, 1: 1e-6, 2: 1e-5, 3: 1e-5, 4: 1e-5}


def load_f64_trace(kind):
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"f64_trace_{F64_TRACES[kind]}.npz")
    return dict(np.load(path))


def check_f64_trace(kind, tr, obs, reward, flags):
    """obs [OD, T], reward [T], flags [T] of ONE step taken from every recorded (state, count, action) of the trace,
    against the float64 answers: |got - want| <= tol * max(1, |want|), tol = 1e-6 (CartPole, MountainCar) or 1e-5;
    flags exact except on marked seams."""
    want_obs, want_rew = tr["obs"].astype(np.float64).T, tr["reward"].astype(np.float64)
    ok = tr["seam"] == 0
    err_obs = np.abs(obs.astype(np.float64) - want_obs) / np.maximum(1.0, np.abs(want_obs))
    err_rew = np.abs(reward.astype(np.float64) - want_rew) / np.maximum(1.0, np.abs(want_rew))
    want_flags = (tr["terminated"] | (tr["truncated"] << 1)).astype(np.uint8)
    assert err_obs[:, ok].max() <= F64_TOL[kind], (KIND_NAMES[kind], "obs", float(err_obs[:, ok].max()))
    assert err_rew[ok].max() <= F64_TOL[kind], (KIND_NAMES[kind], "reward", float(err_rew[ok].max()))
    assert np.array_equal(flags[ok], want_flags[ok]), (KIND_NAMES[kind], "flags", int((flags[ok] != want_flags[ok]).sum()))
    return float(err_obs[:, ok].max()), float(err_rew[ok].max())
