"""Step counters as START STAMPS (CounterMode in mgym_kernels.cuh): an auto-reset handle keeps, per env, the step
index at which the episode began; a step only reads it (or, without a time limit, does not touch it until the env
finishes).  Everything observable -- truncation, get_state counts, episode-length statistics -- must equal the
oracle's plain counters (cartpole.rs:296-306 `steps_since_reset`), also across the 16-bit wrap of the stamp."""
import numpy as np
import pytest

from helpers import KIND_NAMES, assert_bit_equal, random_actions, random_states

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


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


def check_counts(env, ref, what):
    state, steps, _ = env.get_state()
    assert_bit_equal(host(state), ref.state, what + ": state")
    assert_bit_equal(host(steps).view(np.uint32), ref.steps, what + ": step counts")
    s = env.stats()
    assert (s.episodes, s.terminated, s.truncated, s.length_sum) == (
        ref.stats.episodes, ref.stats.terminated, ref.stats.truncated, ref.stats.length_sum), what


@pytest.mark.parametrize("kind", [0, 3])
def test_sixteen_bit_stamps_wrap(gym, oracle, kind):
    """CartPole (500) and Pendulum (200) keep 16-bit stamps: drive the handle's step index across 2^16 with fused
    rollouts, then cross the wrap with per-call steps; flags, counts and statistics follow the oracle throughout."""
    n, seed = 1024, 0xC0FFEE
    env = gym.GpuVecEnv(kind, n, seed=seed)
    ref = oracle.VecState(kind, n, auto_reset=1, seed=seed)
    assert_bit_equal(host(env.reset()), ref.reset(), "reset")
    K = 6550
    for chunk in range(10):  # device policy: 65500 steps
        out = env.rollout(K, want_obs=False)
        _, _, f, dc = ref.rollout(K, None)
        assert_bit_equal(host(out.flags), f, f"flags of chunk {chunk}")
        assert int(out.done_count.item()) == dc
    assert env.step_index == ref.t == 65500
    check_counts(env, ref, "before the wrap")
    rng = np.random.default_rng(5)
    for t in range(80):  # t = 65500 .. 65579 crosses 65536
        a = random_actions(rng, kind, n)
        info = env.step(dev(a))
        o, r, f = ref.step(a)
        assert_bit_equal(host(info.flags), f, f"flags at t={65500 + t}")
        assert_bit_equal(host(info.state), o, f"obs at t={65500 + t}")
    check_counts(env, ref, "after the wrap")
    out = env.rollout(700)  # every env truncates at least once more
    o, r, f, dc = ref.rollout(700, None)
    assert_bit_equal(host(out.flags), f, "flags after the wrap")
    assert_bit_equal(host(out.obs), o, "obs after the wrap")
    check_counts(env, ref, "end")
    env.close()


@pytest.mark.parametrize("n", [2048, 3076, 1027])
def test_lazy_stamps_mountain_car(gym, oracle, n):
    """MountainCar-v0 with statistics but no time limit (the default handle) moves no counter in the per-call step:
    the stamp is read only when an env reaches the goal.  Mix per-call steps, rollouts, masked resets and injected
    counts; the episode lengths the statistics see must be the oracle's."""
    seed = 77
    env = gym.GpuVecEnv(1, n, seed=seed)
    ref = oracle.VecState(1, n, auto_reset=1, seed=seed)
    rng = np.random.default_rng(n)
    assert_bit_equal(host(env.reset()), ref.reset(), "reset")
    st = random_states(rng, 1, n)
    counts = rng.integers(0, 5000, n).astype(np.uint32)
    env.set_state(dev(st), dev(counts.view(np.int32)))
    ref.state[:], ref.steps[:] = st, counts
    push_right = np.full(n, 2, np.uint8)
    for phase in range(3):
        for t in range(40):
            a = push_right if t % 3 else random_actions(rng, 1, n)
            info = env.step(dev(a))
            o, r, f = ref.step(a)
            assert_bit_equal(host(info.flags), f, f"flags phase {phase} t={t}")
            assert_bit_equal(host(info.state), o, f"obs phase {phase} t={t}")
        check_counts(env, ref, f"steps of phase {phase}")
        a = np.stack([push_right if t % 2 else random_actions(rng, 1, n) for t in range(60)])
        out = env.rollout(60, dev(a))
        o, r, f, dc = ref.rollout(60, a)
        assert_bit_equal(host(out.flags), f, f"rollout flags phase {phase}")
        assert_bit_equal(host(out.obs), o, f"rollout obs phase {phase}")
        check_counts(env, ref, f"rollout of phase {phase}")
        mask = (rng.random(n) < 0.25).astype(np.uint8)
        assert_bit_equal(host(env.reset(mask=dev(mask))), ref.reset(mask=mask), f"masked reset phase {phase}")
        check_counts(env, ref, f"masked reset of phase {phase}")
    assert ref.stats.episodes > n // 16  # the envs placed next to the goal, at least
    blob = env.checkpoint()
    other = gym.GpuVecEnv(1, n, seed=1)
    other.restore(blob)
    for e in (env, other):
        e.step(dev(push_right))
    ref.step(push_right)
    check_counts(env, ref, "after the checkpoint")
    check_counts(other, ref, "restored handle")
    env.close(), other.close()


@pytest.mark.parametrize("kind", [2, 3])
def test_returns_are_opt_in(gym, oracle, kind):
    """MountainCarContinuous / Pendulum: the per-env running return (8 bytes per env-step in the step kernel) is only
    kept with track_returns; without it everything else is unchanged and return_sum reads 0."""
    n, T, seed = 2048, 260, 3
    lean = gym.GpuVecEnv(kind, n, seed=seed)
    full = gym.GpuVecEnv(kind, n, seed=seed, track_returns=True)
    ref = oracle.VecState(kind, n, auto_reset=1, seed=seed)
    rng = np.random.default_rng(kind)
    ref.reset()
    for e in (lean, full):
        e.reset()
    st = random_states(rng, kind, n)
    for e in (lean, full):
        e.set_state(dev(st))
    ref.state[:] = st
    for t in range(T):
        a = random_actions(rng, kind, n)
        o, r, f = ref.step(a)
        for e in (lean, full):
            info = e.step(dev(a))
            assert_bit_equal(host(info.flags), f, f"{KIND_NAMES[kind]} flags t={t}")
            assert_bit_equal(host(info.state), o, f"{KIND_NAMES[kind]} obs t={t}")
            assert_bit_equal(host(info.reward), r, f"{KIND_NAMES[kind]} reward t={t}")
    check_counts(lean, ref, "lean")
    check_counts(full, ref, "full")
    assert ref.stats.episodes > 0
    assert lean.stats().return_sum == 0.0
    assert full.stats().return_sum == pytest.approx(ref.stats.return_sum, rel=1e-9, abs=1e-6)
    lean.close(), full.close()


def test_checkpoint_rejects_a_different_configuration(gym):
    """A blob only loads into a handle whose dynamics-relevant configuration matches: the
    continuation would otherwise silently diverge."""
    n = 1024
    a = gym.GpuVecEnv(1, n, seed=5, goal_velocity=0.02)
    a.reset()
    a.rollout(10)
    blob = a.checkpoint()
    assert blob == a.checkpoint()  # no uninitialised padding: byte-for-byte reproducible
    gym.GpuVecEnv(1, n, seed=9, goal_velocity=0.02).restore(blob)  # the seed travels with the blob
    for cfg in (dict(), dict(goal_velocity=0.02, max_episode_steps=200), dict(goal_velocity=0.02, env_index_base=4096),
                dict(goal_velocity=0.02, track_stats=False)):
        with pytest.raises(gym.MgymError):
            gym.GpuVecEnv(1, n, seed=5, **cfg).restore(blob)
    b = gym.GpuVecEnv(0, n, seed=5, sutton_barto_reward=True)
    b.reset()
    blob = b.checkpoint()
    for cfg in (dict(), dict(sutton_barto_reward=True, is_euler=False)):
        with pytest.raises(gym.MgymError):
            gym.GpuVecEnv(0, n, seed=5, **cfg).restore(blob)


def test_injected_counts_past_a_sixteen_bit_limit_still_truncate(gym, oracle):
    """mgym_set_state with counts beyond what a 16-bit stamp can hold (CartPole: limit 500): whatever the value, the
    env is past its limit and the next step truncates, exactly as a plain counter would; the handle keeps such a count
    as 65535 (include/mgym.h), which is what the oracle is given here -- the count also keys the reset stream."""
    n = 1024
    env = gym.GpuVecEnv(0, n, seed=3)
    ref = oracle.VecState(0, n, auto_reset=1, seed=3)
    env.reset(), ref.reset()
    counts = np.array([0, 498, 499, 500, 65535, 65536, 65536 + 7, 70000, 1 << 31, 0xFFFFFFFE] * 103, np.uint32)[:n]
    z = np.zeros((4, n), np.float32)
    env.set_state(dev(z), dev(counts.view(np.int32)))
    ref.state[:], ref.steps[:] = z, np.minimum(counts, 65535)
    a = np.zeros(n, np.uint8)
    info = env.step(dev(a))
    o, r, f = ref.step(a)
    assert_bit_equal(host(info.flags), f, "flags")
    assert_bit_equal(host(info.state), o, "obs after the same-step reset")
    assert (f[counts >= 499] == 2).all() and (f[counts < 499] == 0).all()
    s = env.stats()
    assert (s.episodes, s.truncated) == (ref.stats.episodes, ref.stats.truncated)
    env.close()


@pytest.mark.parametrize("n", [1024, 515])
def test_cartpole_rollout_reset_sources(gym, oracle, n):
    """The CartPole rollout draws reset states AHEAD of time, keyed by the start of the running episode
    (rollout_kernel, Env<0>::PREFETCH_RESETS).  Three situations the ordinary parity runs reach only by chance:
    every env of every lane finishing in the very first step of a launch (no state drawn yet, four pending slots
    per lane), an injected reset pool taking precedence over the drawn states, and going back to Philox after the
    pool is removed -- trajectories, counts and statistics must be the oracle's throughout."""
    seed, K = 21, 40
    env = gym.GpuVecEnv(0, n, seed=seed)
    ref = oracle.VecState(0, n, auto_reset=1, seed=seed)
    rng = np.random.default_rng(n)
    assert_bit_equal(host(env.reset()), ref.reset(), "reset")

    def play(what, with_actions):
        a = random_actions(rng, 0, (K, n)) if with_actions else None
        out = env.rollout(K, None if a is None else dev(a))
        o, r, f, dc = ref.rollout(K, a)
        assert_bit_equal(host(out.flags), f, what + ": flags")
        assert_bit_equal(host(out.obs), o, what + ": obs")
        assert int(out.done_count.item()) == dc
        check_counts(env, ref, what)

    fallen = np.zeros((4, n), np.float32)
    fallen[2] = 0.3  # beyond the 12-degree threshold: every env terminates in the first step
    counts = rng.integers(0, 400, n).astype(np.uint32)
    env.set_state(dev(fallen), dev(counts.view(np.int32)))
    ref.state[:], ref.steps[:] = fallen, counts
    play("all envs finish in step 0", True)
    play("steady state, device policy", False)
    pool = random_states(rng, 0, 61)
    pool[2, ::3] = 0.25  # a third of the pool states terminate at once: envs that finish every step
    env.set_reset_pool(dev(pool))
    ref.set_reset_pool(pool)
    play("injected pool", True)
    env.set_reset_pool(None)
    ref.set_reset_pool(None)
    play("back to Philox", True)
    env.close()
