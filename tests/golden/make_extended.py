#!/usr/bin/env python
"""Extended known-answer set: long auto-reset traces of every kind, produced by the CPU oracle
and committed as digests, so that the oracle AND the device path are both held to fixed answers (the dynamic
parity tests alone would not notice the two drifting together).  The traces reach what the reference's 100-step
fixtures never do: CartPole's 500-step truncation, MountainCar's wall and goal, Pendulum/Acrobot time limits.

Run in the build container:  python tests/golden/make_extended.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = [  # kind, envs, steps, config
    (0, 256, 1200, {}),
    (0, 64, 700, {"sutton_barto_reward": 1}),
    (1, 256, 1500, {"max_episode_steps": 200}),
    (1, 256, 1500, {}),
    (2, 256, 1200, {}),
    (3, 256, 450, {}),
    (4, 128, 700, {}),
]
SEED = 0x600D5EED


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_case(o, kind, n, steps, cfg):
    vs = o.VecState(kind, n, auto_reset=1, seed=SEED, **cfg)
    if kind == 0 and not cfg:
        vs.cfg.auto_reset = 1
    first = vs.reset()
    obs, rew, flg, dones = vs.rollout(steps)  # device-policy actions: a pure function of (seed, env, step)
    return {
        "kind": kind, "envs": n, "steps": steps, "config": cfg, "seed": SEED,
        "sha256": digest(first, obs, rew, flg), "done_steps": int(dones),
        "terminated": int(((flg & 1) != 0).sum()), "truncated": int(((flg & 2) != 0).sum()),
        "final_obs_env0": [float(x) for x in obs[-1, :, 0]],
        "stats": [int(vs.stats.episodes), int(vs.stats.terminated), int(vs.stats.truncated), int(vs.stats.length_sum)],
    }


if __name__ == "__main__":
    from oracle import oracle as o

    out = [run_case(o, *c) for c in CASES]
    for r in out:
        print(o.KIND_NAMES[r["kind"]], r["config"], "done", r["done_steps"], "term", r["terminated"], "trunc", r["truncated"])
    with open(os.path.join(HERE, "extended_traces.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_extended.py (CPU oracle)", "cases": out}, f, indent=1)
