"""Committed digests of long auto-reset traces (tests/golden/extended_traces.json): the oracle on CPU and the
CUDA path on GPU must both reproduce them."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
CASES = json.load(open(os.path.join(HERE, "golden", "extended_traces.json")))["cases"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"kind{c['kind']}-{c['steps']}-{'-'.join(c['config']) or 'default'}")
def test_oracle_reproduces_extended_traces(oracle, case):
    import make_extended

    got = make_extended.run_case(oracle, case["kind"], case["envs"], case["steps"], case["config"])
    assert got["sha256"] == case["sha256"]
    assert got["stats"] == case["stats"] and got["done_steps"] == case["done_steps"]


def test_extended_traces_reach_the_untested_branches():
    by = {(c["kind"], json.dumps(c["config"], sort_keys=True)): c for c in CASES}
    assert by[(0, "{}")]["terminated"] > 1000                    # cartpole.rs:319-329
    assert by[(1, '{"max_episode_steps": 200}')]["truncated"] > 1000
    assert by[(3, "{}")]["truncated"] == 256 * 2                 # Pendulum TimeLimit 200, twice in 450 steps
    assert by[(4, "{}")]["terminated"] > 0 and by[(4, "{}")]["truncated"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"kind{c['kind']}-{c['steps']}-{'-'.join(c['config']) or 'default'}")
def test_gpu_reproduces_extended_traces(case):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import make_extended
    import modurl_gym_b200 as m

    cfg = {k: (bool(v) if k == "sutton_barto_reward" else v) for k, v in case["config"].items()}
    env = m.GpuVecEnv(case["kind"], case["envs"], seed=case["seed"], **cfg)
    first = env.reset().cpu().numpy().copy()
    out = env.rollout(case["steps"])
    got = make_extended.digest(first, out.obs.cpu().numpy(), out.reward.cpu().numpy(), out.flags.cpu().numpy())
    assert got == case["sha256"]
    s = env.stats()
    assert [s.episodes, s.terminated, s.truncated, s.length_sum] == case["stats"]
    assert int(out.done_count.item()) == case["done_steps"]
