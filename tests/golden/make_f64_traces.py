#!/usr/bin/env python
"""10^4-step teacher-forced known-answer traces of all five kinds, computed by the independent float64 transcription
in tests/f64_gymnasium.py -- NOT by the oracle or the device code: the three kinds the reference does not implement
(MountainCarContinuous-v0, Pendulum-v1, Acrobot-v1; Gymnasium's equations) and CartPole-v1 / MountainCar-v0 (the
reference's equations), whose traces reach the 500-step truncation, the wall and the goal that the reference's
100-step fixtures never do.  Gymnasium itself is not
installed in this image, so these are the closest thing to an external answer this image allows:
SURVEY rows A6-A8 stay "parity unpinned", but north_star's 1e-5 bar is exercised over 10^4-step traces.

Format (npz, answers rounded to f32: 6e-8 relative, far below the 1e-5 bar): per step t the f32 input `state`, the
episode step `count`, the `action`, and the answer `obs`, `reward`, `terminated`, `truncated`, `seam` (1 = a clamp /
wrap / threshold decided within rounding distance: flags and the affected components are not compared there).

Run in the build container:  python tests/golden/make_f64_traces.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import f64_gymnasium as g  # noqa: E402

STEPS = 10_000
CASES = {"cartpole": (g.CARTPOLE, 0xA2), "mountain_car": (g.MOUNTAIN_CAR, 0xA5),
         "mountain_car_continuous": (g.MOUNTAIN_CAR_CONTINUOUS, 0xA6), "pendulum": (g.PENDULUM, 0xA7),
         "acrobot": (g.ACROBOT, 0xA8)}

if __name__ == "__main__":
    for name, (kind, seed) in CASES.items():
        tr = g.teacher_forced_trace(kind, STEPS, seed)
        path = os.path.join(HERE, f"f64_trace_{name}.npz")
        np.savez_compressed(path, state=tr["state"], count=tr["count"], action=tr["action"],
                            obs=tr["obs"].astype(np.float32), reward=tr["reward"].astype(np.float32),
                            terminated=tr["terminated"], truncated=tr["truncated"], seam=tr["seam"])
        print(f"{name}: {STEPS} steps, terminated {int(tr['terminated'].sum())}, truncated {int(tr['truncated'].sum())}, "
              f"seam {int(tr['seam'].sum())}, {os.path.getsize(path) / 1024:.0f} KiB")
