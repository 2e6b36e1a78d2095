"""Randomised differential sessions: a seeded generator draws a handle configuration (kind, size, auto/manual,
options, index base, time limit) and a sequence of calls (step, fused rollout with given or device-drawn actions,
full and masked resets, state injection); the CUDA path and the CPU oracle run the same session and every output
must agree bit for bit.  Complements the targeted tests of test_gpu_parity.py with combinations nobody wrote down."""
import numpy as np
import pytest

from helpers import KIND_NAMES, assert_bit_equal, random_actions, random_states

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gym():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import modurl_gym_b200 as m

    m.load_library()
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


SIZES = [1, 7, 128, 515, 1024, 1027, 2048, 3076, 4100, 5120]


@pytest.mark.parametrize("session", range(24))
def test_random_session_matches_oracle(gym, oracle, session):
    rng = np.random.default_rng(90_000 + session)
    kind = int(rng.integers(0, 5))
    n = int(rng.choice(SIZES))
    auto = bool(rng.integers(0, 2))
    cfg = {}
    if kind == 0 and rng.random() < 0.3:
        cfg["sutton_barto_reward"] = True
    if kind == 1 and rng.random() < 0.5:
        cfg["goal_velocity"] = float(rng.choice([0.0, 0.01, 0.03]))
    if kind != 0 and rng.random() < 0.6:
        cfg["max_episode_steps"] = int(rng.choice([3, 17, 64, 200]))
    if rng.random() < 0.5:
        cfg["env_index_base"] = int(rng.choice([4, 4096, 1 << 33]))
    seed = int(rng.integers(0, 1 << 40))
    ocfg = {k: (int(v) if k != "goal_velocity" else v) for k, v in cfg.items()}
    env = gym.GpuVecEnv(kind, n, auto_reset=auto, seed=seed, track_returns=True, **cfg)
    ref = oracle.VecState(kind, n, auto_reset=int(auto), seed=seed, **ocfg)
    what = f"session {session}: {KIND_NAMES[kind]} n={n} auto={auto} {cfg}"
    assert_bit_equal(host(env.reset()), ref.reset(), what + " reset")
    last_flags = np.zeros(n, np.uint8)
    for call in range(int(rng.integers(6, 14))):
        op = rng.choice(["step", "step", "step", "rollout", "rollout_policy", "reset", "masked_reset", "set_state"])
        tag = f"{what} call {call} ({op})"
        if op == "step":
            for _ in range(int(rng.integers(1, 12))):
                a = random_actions(rng, kind, n)
                info = env.step(dev(a))
                o, r, f = ref.step(a)
                assert_bit_equal(host(info.flags), f, tag + " flags")
                assert_bit_equal(host(info.state), o, tag + " obs")
                assert_bit_equal(host(info.reward), r, tag + " reward")
                last_flags = f
        elif op in ("rollout", "rollout_policy"):
            K = int(rng.integers(1, 20))
            a = None if op == "rollout_policy" else random_actions(rng, kind, (K, n))
            out = env.rollout(K, None if a is None else dev(a))
            o, r, f, dc = ref.rollout(K, a)
            assert_bit_equal(host(out.flags), f, tag + " flags")
            assert_bit_equal(host(out.obs), o, tag + " obs")
            assert_bit_equal(host(out.reward), r, tag + " reward")
            assert int(out.done_count.item()) == dc, tag
            last_flags = f[-1]
        elif op == "reset":
            assert_bit_equal(host(env.reset()), ref.reset(), tag)
        elif op == "masked_reset":
            mask = (last_flags != 0).astype(np.uint8) if rng.random() < 0.5 else (rng.random(n) < 0.3).astype(np.uint8)
            assert_bit_equal(host(env.reset(mask=dev(mask))), ref.reset(mask=mask), tag)
        else:
            st = random_states(rng, kind, n)
            env.set_state(dev(st))
            ref.state[:, :n] = st
            ref.steps[:] = 0
            ref.sbt[:] = 0
            ref.ep_return[:] = 0
    state, steps, sbt = env.get_state()
    assert_bit_equal(host(state), ref.state, what + " final state")
    assert_bit_equal(host(steps).view(np.uint32), ref.steps, what + " final step counters")
    if auto:
        s = env.stats()
        assert (s.episodes, s.terminated, s.truncated, s.length_sum) == (
            ref.stats.episodes, ref.stats.terminated, ref.stats.truncated, ref.stats.length_sum), what
        assert s.return_sum == pytest.approx(ref.stats.return_sum, rel=1e-9, abs=1e-6)
    assert env.step_index == ref.t, what
    env.close()
