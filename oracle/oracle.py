"""ctypes loader for the CPU oracle (oracle/mgym_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (modurl_gym_b200) never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgym_oracle.so")

CARTPOLE, MOUNTAIN_CAR, MOUNTAIN_CAR_CONTINUOUS, PENDULUM, ACROBOT = range(5)
KIND_NAMES = ["CartPole-v1", "MountainCar-v0", "MountainCarContinuous-v0", "Pendulum-v1", "Acrobot-v1"]
STATE_DIM = [4, 2, 2, 2, 4]
OBS_DIM = [4, 2, 2, 3, 6]
CONTINUOUS = [False, False, True, True, False]
FLAG_TERMINATED, FLAG_TRUNCATED = 1, 2


class Config(C.Structure):
    _fields_ = [
        ("auto_reset", C.c_int32),
        ("max_episode_steps", C.c_int32),
        ("sutton_barto_reward", C.c_int32),
        ("is_euler", C.c_int32),
        ("goal_velocity", C.c_float),
        ("seed", C.c_uint64),
        ("env_index_base", C.c_uint64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("episodes", C.c_uint64),
        ("terminated", C.c_uint64),
        ("truncated", C.c_uint64),
        ("length_sum", C.c_uint64),
        ("return_sum", C.c_double),
    ]


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("mgym_oracle.c", "mgym_oracle.h", "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.run(["make", "-C", HERE, "libmgym_oracle.so", "exhaustive_trig"], check=True,
                       capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_sinf.restype = C.c_float
        L.oracle_sinf.argtypes = [C.c_float]
        L.oracle_cosf.restype = C.c_float
        L.oracle_cosf.argtypes = [C.c_float]
        L.oracle_trig_checksum.restype = C.c_uint64
        L.oracle_trig_checksum.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_int]
        L.oracle_config_default.argtypes = [C.c_int, C.POINTER(Config)]
        L.oracle_env_step.restype = C.c_uint32
        L.oracle_env_step.argtypes = [C.c_int, C.POINTER(Config), C.c_void_p, C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint32), C.c_uint32, C.c_float, C.POINTER(C.c_float)]
        L.oracle_env_obs.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_reset_state.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.oracle_sample_action.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.oracle_vec_step.argtypes = [C.c_int, C.POINTER(Config), C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_void_p] * 6 + [
            C.c_uint64] + [C.c_void_p] * 4 + [C.POINTER(Stats)]
        L.oracle_vec_rollout.argtypes = [C.c_int, C.POINTER(Config), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32] + [
            C.c_void_p] * 6 + [C.c_uint64] + [C.c_void_p] * 3 + [C.POINTER(C.c_uint64), C.POINTER(Stats)]
        L.oracle_vec_reset.argtypes = [C.c_int, C.POINTER(Config), C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_void_p] * 6 + [
            C.c_uint64, C.c_void_p]
        L.oracle_baseline_loop.restype = C.c_uint64
        L.oracle_baseline_loop.argtypes = [C.c_int, C.POINTER(Config), C.c_uint64, C.c_int,
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def default_config(kind, **over):
    cfg = Config()
    lib().oracle_config_default(kind, C.byref(cfg))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sinf(x):
    return lib().oracle_sinf(float(np.float32(x)))


def cosf(x):
    return lib().oracle_cosf(float(np.float32(x)))


def trig_checksum(first_bits, count, stride=1, threads=None):
    return lib().oracle_trig_checksum(first_bits, count, stride, threads or (os.cpu_count() or 1))


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def reset_state(kind, seed, g, t, tag):
    st = np.zeros(4, dtype=np.float32)
    lib().oracle_reset_state(kind, seed, g, t, tag, _p(st))
    return st[: STATE_DIM[kind]]


def sample_action(kind, seed, g, t):
    a8 = np.zeros(1, dtype=np.uint8)
    af = np.zeros(1, dtype=np.float32)
    lib().oracle_sample_action(kind, seed, g, t, _p(a8), _p(af))
    return af[0] if CONTINUOUS[kind] else a8[0]


class ScalarEnv:
    """One env with the reference's Gym semantics (reset/step -> obs, reward, done, truncated).

    Mirrors `impl Gym for CartPoleV1` (cartpole.rs:234-357) / `MountainCarV0`
    (mountain_car.rs:275-339) plus the `Testable` hooks (cartpole.rs:436-447)."""

    def __init__(self, kind, **cfg):
        self.kind = kind
        self.cfg = default_config(kind, **cfg)
        self.state = np.zeros(4, dtype=np.float32)  # cartpole.rs:85 zero state
        self.steps = C.c_uint32(0)
        self.sbt = C.c_uint32(1 if kind == CARTPOLE else 0)  # cartpole.rs:81 Some(0)
        self.n_resets = 0
        self.index = 0

    def reset(self):
        st = reset_state(self.kind, self.cfg.seed, self.cfg.env_index_base + self.index, self.n_resets, 1)
        self.n_resets += 1
        self.state[:] = 0
        self.state[: len(st)] = st
        self.steps.value = 0
        self.sbt.value = 0
        return self.obs()

    def reset_deterministic(self):  # cartpole.rs:437-442, mountain_car.rs:403-408
        if self.kind == CARTPOLE:
            self.reset()
        self.state[:] = 0
        return self.obs()

    def set_state(self, state):  # cartpole.rs:444-446
        self.state[:] = 0
        self.state[: len(state)] = np.asarray(state, dtype=np.float32)

    def obs(self):
        o = np.zeros(6, dtype=np.float32)
        lib().oracle_env_obs(self.kind, _p(self.state), _p(o))
        return o[: OBS_DIM[self.kind]].copy()

    def step(self, action):
        r = C.c_float(0)
        au = 0 if CONTINUOUS[self.kind] else int(action)
        af = float(action) if CONTINUOUS[self.kind] else 0.0
        flags = lib().oracle_env_step(self.kind, C.byref(self.cfg), _p(self.state), C.byref(self.steps),
                                      C.byref(self.sbt), au, af, C.byref(r))
        return self.obs(), r.value, bool(flags & FLAG_TERMINATED), bool(flags & FLAG_TRUNCATED)


class VecState:
    """Host-side SoA env state for the batched oracle drivers."""

    def __init__(self, kind, n, ld=None, **cfg):
        self.kind, self.n = kind, n
        self.ld = ld or n
        self.cfg = default_config(kind, **cfg)
        self.state = np.zeros((STATE_DIM[kind], self.ld), dtype=np.float32)
        self.steps = np.zeros(self.ld, dtype=np.uint32)
        self.sbt = np.full(self.ld, 1 if kind == CARTPOLE else 0, dtype=np.uint32)
        self.ep_return = np.zeros(self.ld, dtype=np.float32)
        self.stats = Stats()
        self.t = 0
        self.n_resets = 0
        self.reset_pool = None

    def set_reset_pool(self, pool):
        self.reset_pool = None if pool is None else np.ascontiguousarray(pool, dtype=np.float32)

    def _pool(self):
        if self.reset_pool is None:
            return None, 0
        return _p(self.reset_pool), self.reset_pool.shape[1]

    def reset(self, mask=None):
        obs = np.zeros((OBS_DIM[self.kind], self.ld), dtype=np.float32)
        pp, pl = self._pool()
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().oracle_vec_reset(self.kind, C.byref(self.cfg), self.n, self.ld, self.n_resets, _p(m), _p(self.state),
                               _p(self.steps), _p(self.sbt), _p(self.ep_return), pp, pl, _p(obs))
        self.n_resets += 1
        return obs

    def step(self, actions, want_final_obs=False):
        od = OBS_DIM[self.kind]
        obs = np.zeros((od, self.ld), dtype=np.float32)
        rew = np.zeros(self.ld, dtype=np.float32)
        flg = np.zeros(self.ld, dtype=np.uint8)
        fin = np.zeros((od, self.ld), dtype=np.float32) if want_final_obs else None
        a = np.ascontiguousarray(actions, dtype=np.float32 if CONTINUOUS[self.kind] else np.uint8)
        pp, pl = self._pool()
        lib().oracle_vec_step(self.kind, C.byref(self.cfg), self.n, self.ld, self.t, _p(self.state), _p(self.steps),
                              _p(self.sbt), _p(self.ep_return), _p(a), pp, pl, _p(obs), _p(rew), _p(flg), _p(fin),
                              C.byref(self.stats))
        self.t += 1
        return (obs, rew, flg, fin) if want_final_obs else (obs, rew, flg)

    def rollout(self, K, actions=None):
        od = OBS_DIM[self.kind]
        obs = np.zeros((K, od, self.ld), dtype=np.float32)
        rew = np.zeros((K, self.ld), dtype=np.float32)
        flg = np.zeros((K, self.ld), dtype=np.uint8)
        a = None
        if actions is not None:
            a = np.ascontiguousarray(actions, dtype=np.float32 if CONTINUOUS[self.kind] else np.uint8)
        dc = C.c_uint64(0)
        pp, pl = self._pool()
        lib().oracle_vec_rollout(self.kind, C.byref(self.cfg), self.n, self.ld, self.t, K, _p(self.state), _p(self.steps),
                                 _p(self.sbt), _p(self.ep_return), _p(a), pp, pl, _p(obs), _p(rew), _p(flg),
                                 C.byref(dc), C.byref(self.stats))
        self.t += K
        return obs, rew, flg, dc.value


def baseline_loop(kind, steps_per_thread, n_threads, **cfg):
    c = default_config(kind, **cfg)
    sec = C.c_double(0)
    chk = C.c_double(0)
    total = lib().oracle_baseline_loop(kind, C.byref(c), steps_per_thread, n_threads, C.byref(sec), C.byref(chk))
    return total, sec.value, chk.value
