"""The f32 oracle against 10^4-step teacher-forced traces computed by the independent float64 transcription
(tests/f64_gymnasium.py): the three kinds the reference does not implement (parity unpinned) held
to north_star's 1e-5 bar over whole traces, time limits and terminations included; and CartPole / MountainCar (the
reference's equations, 1e-6) through the 500-step truncation, the wall and the goal that the reference's own 100-step
fixtures never reach.  Every
step of a trace is an independent (state, count, action) -> answer record (src/testing.rs:65-134 protocol), so the
whole trace is replayed as ONE manual-mode step of a 10^4-env batch."""
import numpy as np
import pytest

from helpers import F64_TRACES, check_f64_trace, load_f64_trace


def test_fixture_is_what_the_generator_produces():
    """The committed fixture is reproducible from the committed generator (first 600 steps of each trace)."""
    import f64_gymnasium as g
    from golden.make_f64_traces import CASES

    for name, (kind, seed) in CASES.items():
        tr, fresh = load_f64_trace(kind), g.teacher_forced_trace(kind, 600, seed)
        assert np.array_equal(tr["state"][:600], fresh["state"]) and np.array_equal(tr["count"][:600], fresh["count"])
        assert np.array_equal(tr["action"][:600], fresh["action"])
        assert np.array_equal(tr["obs"][:600], fresh["obs"].astype(np.float32))
        assert np.array_equal(tr["terminated"][:600], fresh["terminated"])


@pytest.mark.parametrize("kind", sorted(F64_TRACES))
def test_oracle_follows_the_float64_traces(oracle, kind):
    tr = load_f64_trace(kind)
    T = tr["state"].shape[0]
    assert T == 10_000 and (kind == 1 or tr["truncated"].sum() >= 9) and (kind == 3 or tr["terminated"].sum() >= 9)
    if kind == 1:
        assert (tr["obs"][:, 0] == np.float32(-1.2)).sum() >= 5  # the wall rule (mountain_car.rs:311-313) is reached
    ref = oracle.VecState(kind, T, auto_reset=0)
    ref.state[:] = tr["state"].T
    ref.steps[:] = tr["count"]
    ref.sbt[:] = 0  # steps_beyond_terminated = None: every record is a step of a running episode (cartpole.rs:239)
    obs, rew, flg = ref.step(np.ascontiguousarray(tr["action"]))
    eo, er = check_f64_trace(kind, tr, obs, rew, flg)
    print(f"kind {kind}: max relative error obs {eo:.2e} reward {er:.2e}")
