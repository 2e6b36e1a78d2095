"""Pins the CPU oracle against the reference's own known-answer fixtures.

Replays python_tests/{cartpole,mountain_car} (copied to tests/golden/ by
tests/golden/make_golden.py) with the teacher-forced protocol of
/root/reference/src/testing.rs:65-134.  The reference accepts 1e-4 absolute; the f32
restatement is held to 2e-7 (the fixtures are float64 Gymnasium results cast to f32, so
they pin values to ~1 ulp, not bitwise -- SURVEY.md section 4)."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF_TOL = 1e-4   # src/testing.rs:42-45
TIGHT_TOL = 2e-7


def load(name):
    with open(os.path.join(GOLDEN, f"{name}_gymnasium.json")) as f:
        return json.load(f)


def replay(env, doc, tol):
    """src/testing.rs:65-134, with `env` playing the role of `T: Gym + Testable`."""
    actions, exp_obs = doc["actions"], doc["observation"]
    env.reset_deterministic()                                   # :65
    worst = 0.0
    for i, action in enumerate(actions):
        if i == 0 or doc["done"][i - 1]:                        # :73-75
            env.reset_deterministic()
        else:                                                   # :77-86 teacher forcing
            env.set_state(exp_obs[i - 1])
        obs, reward, done, truncated = env.step(action)         # :89
        assert abs(reward - doc["reward"][i]) <= tol, (i, reward)      # :99
        assert done == doc["done"][i], f"done mismatch at step {i + 1}"            # :106
        assert truncated == doc["truncated"][i], f"truncated mismatch at step {i + 1}"  # :114
        err = np.max(np.abs(obs.astype(np.float64) - np.asarray(exp_obs[i])))
        assert err < tol, (i, obs, exp_obs[i])                  # :124-133
        worst = max(worst, float(err))
    return worst


@pytest.mark.parametrize("name,kind", [("cartpole", 0), ("mountain_car", 1)])
def test_oracle_against_python(oracle, name, kind):
    doc = load(name)
    assert len(doc["actions"]) == len(doc["observation"]) == 100
    worst = replay(oracle.ScalarEnv(kind), doc, TIGHT_TOL)
    assert worst < TIGHT_TOL < REF_TOL


def test_fixture_facts():
    """Facts SURVEY.md section 4 records about the fixtures; guards against a bad copy."""
    cp, mc = load("cartpole"), load("mountain_car")
    assert [i for i, d in enumerate(cp["done"]) if d] == [27, 54, 96]
    assert not any(cp["truncated"]) and set(cp["reward"]) == {1.0} and set(cp["actions"]) <= {0, 1}
    assert not any(mc["done"]) and set(mc["reward"]) == {-1.0} and set(mc["actions"]) <= {0, 1, 2}


# --- behavioural pins from the reference's unit tests ---------------------------------

def test_cartpole(oracle):
    """cartpole.rs:365-390 test_cartpole"""
    env = oracle.ScalarEnv(oracle.CARTPOLE)
    state = env.reset()
    assert state.shape == (4,)
    assert np.all(np.abs(state) <= 0.05)
    obs, reward, done, _ = env.step(0)
    assert obs.shape == (4,) and reward == 1.0 and not done


def test_reward_is_one_when_not_terminated(oracle):
    """cartpole.rs:405-434: constant right push terminates within 1 + 50 steps"""
    env = oracle.ScalarEnv(oracle.CARTPOLE)
    env.reset()
    _, reward, done, _ = env.step(1)
    assert reward == 1.0 and not done
    for _ in range(50):
        _, _, done, _ = env.step(1)
        if done:
            break
    assert done


def test_mountain_car(oracle):
    """mountain_car.rs:347-372, :387-400"""
    env = oracle.ScalarEnv(oracle.MOUNTAIN_CAR)
    state = env.reset()
    assert state.shape == (2,) and -0.6 <= state[0] <= -0.4 and state[1] == 0.0
    obs, reward, done, trunc = env.step(0)
    assert obs.shape == (2,) and reward == -1.0 and not done and not trunc
