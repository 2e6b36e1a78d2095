"""Space metadata mirroring modurl::spaces::{Discrete, BoxSpace} as the reference uses them
(cartpole.rs:58-69, mountain_car.rs:42-48): `contains`, `sample`, low/high."""
import ctypes as C

import numpy as np

from . import _lib


class Discrete:
    def __init__(self, n):
        self.n = int(n)

    def contains(self, action):
        """Discrete::contains: a rank-0 unsigned value in [0, n) (cartpole.rs:252, :392-403)."""
        a = np.asarray(action)
        return a.ndim == 0 and a.dtype.kind in "ui" and 0 <= int(a) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class BoxSpace:
    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)

    @property
    def shape(self):
        return self.low.shape

    def contains(self, x):
        x = np.asarray(x, dtype=np.float32)
        return x.shape == self.low.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"BoxSpace(low={self.low.tolist()}, high={self.high.tolist()})"


def observation_space(kind):
    lib = _lib.load()
    d = lib.mgym_obs_dim(kind)
    lo, hi = np.zeros(d, np.float32), np.zeros(d, np.float32)
    _lib.check(lib.mgym_space_observation(kind, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)))
    return BoxSpace(lo, hi)


def action_space(kind):
    lib = _lib.load()
    if lib.mgym_action_is_continuous(kind):
        lo, hi = np.zeros(1, np.float32), np.zeros(1, np.float32)
        _lib.check(lib.mgym_space_action(kind, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)))
        return BoxSpace(lo, hi)
    return Discrete(lib.mgym_num_actions(kind))
