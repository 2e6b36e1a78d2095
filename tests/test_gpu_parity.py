"""Parity of the CUDA path (through the C ABI, via GpuVecEnv) against the CPU oracle.

Bar (BASELINE.json north_star): terminated/truncated flags and step counts bit-exact; states and
rewards within 1e-6 relative for CartPole/MountainCar and 1e-5 for the others.  Because the
device code restates glibc's sinf/cosf exactly and never fuses a multiply-add, the tests below
hold every value to BIT equality, which implies both tolerances."""
import json
import os

import numpy as np
import pytest

from helpers import (CONTINUOUS, KIND_NAMES, NUM_ACTIONS, OBS_DIM, STATE_DIM, assert_bit_equal, random_actions,
                     random_states)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = {0: 1e-6, 1: 1e-6, 2: 1e-5, 3: 1e-5, 4: 1e-5}  # north_star tolerances (documented; we get 0)


@pytest.fixture(scope="module")
def gym():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import modurl_gym_b200 as m

    m.load_library()  # fails loudly if libmgym.so is missing
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------
# building blocks: sin/cos and Philox on the device
# ---------------------------------------------------------------------------------------------
def test_device_trig_bit_exact(gym, oracle):
    import ctypes as C

    lib = gym.load_library()
    rng = np.random.default_rng(7)
    xs = np.concatenate([
        rng.uniform(-0.3, 0.3, 200000), rng.uniform(-4, 4, 200000), rng.uniform(-130, 130, 200000),
        rng.standard_normal(100000) * 1e4, rng.standard_normal(50000) * 1e30,
        np.array([0.0, -0.0, 2.0 ** -12, -(2.0 ** -12), np.nextafter(np.float32(2.0 ** -12), 0), 0.7853982, 0.78539819,
                  -0.7853982, 120.0, 119.99999, -120.0, 1e-40, -1e-40, 3.4e38, -3.4e38, np.inf, -np.inf, np.nan]),
        np.arange(-200, 200) * (np.pi / 4),
    ]).astype(np.float32)
    # the bit patterns around the CartPole and MountainCar working ranges, densely
    xs = np.concatenate([xs, np.arange(0x3e567750 - 50000, 0x3e567750 + 50000, dtype=np.uint32).view(np.float32)])
    x = dev(xs)
    outs = [torch.empty_like(x) for _ in range(4)]
    rc = lib.mgym_probe_trig(C.c_void_p(x.data_ptr()), *[C.c_void_p(o.data_ptr()) for o in outs], x.numel(), None)
    assert rc == 0
    torch.cuda.synchronize()
    s, c, s_only, c_only = [host(o) for o in outs]
    ws = np.array([oracle.sinf(v) for v in xs], dtype=np.float32)
    wc = np.array([oracle.cosf(v) for v in xs], dtype=np.float32)
    finite = np.isfinite(xs)
    assert_bit_equal(s[finite], ws[finite], "sincos.sin")
    assert_bit_equal(c[finite], wc[finite], "sincos.cos")
    assert_bit_equal(s_only[finite], ws[finite], "sin_ref")
    assert_bit_equal(c_only[finite], wc[finite], "cos_ref")
    assert np.isnan(s[~finite]).all() and np.isnan(c[~finite]).all()


def test_device_philox_known_answers(gym, oracle):
    import ctypes as C

    lib = gym.load_library()
    rng = np.random.default_rng(3)
    ck = rng.integers(0, 2 ** 32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    ck[0] = 0
    ck[1] = 0xFFFFFFFF
    ck[2] = [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0]
    d = dev(ck)
    out = torch.empty((4096, 4), dtype=torch.int32, device="cuda")
    assert lib.mgym_probe_philox(C.c_void_p(d.data_ptr()), C.c_void_p(out.data_ptr()), 4096, None) == 0
    got = host(out).view(np.uint32)
    # Random123 known-answer vectors
    assert got[0].tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert got[1].tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert got[2].tolist() == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    want = np.stack([oracle.philox(r[:4], r[4:]) for r in ck[:512]])
    assert_bit_equal(got[:512], want, "philox")


# ---------------------------------------------------------------------------------------------
# the reference's own known-answer test, replayed on the GPU (src/testing.rs:34-146)
# ---------------------------------------------------------------------------------------------
def load_golden(name):
    with open(os.path.join(GOLDEN, f"{name}_gymnasium.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name,kind", [("cartpole", 0), ("mountain_car", 1)])
def test_gym_against_python_scalar(gym, name, kind):
    """One env, 100 teacher-forced steps, exactly the loop of testing.rs:65-134."""
    doc = load_golden(name)
    env = gym.GpuVecEnv(kind, 1, auto_reset=False)
    sd = STATE_DIM[kind]
    zeros = torch.zeros((sd, 1))
    env.reset()                      # reset_deterministic(): reset() then zero state (cartpole.rs:437-442)
    env.set_state(zeros)
    for i, action in enumerate(doc["actions"]):
        if i == 0 or doc["done"][i - 1]:
            env.reset()
            env.set_state(zeros)
        else:
            # set_state only overwrites the state tensor (cartpole.rs:444-446): keep the counters
            _, steps, sbt = env.get_state()
            env.set_state(torch.tensor(doc["observation"][i - 1], dtype=torch.float32).reshape(sd, 1), steps, sbt)
        info = env.step(torch.tensor([action], dtype=torch.uint8, device="cuda"))
        obs = host(info.state)[:, 0]
        assert abs(float(info.reward[0]) - doc["reward"][i]) <= 1e-4                 # testing.rs:99
        assert bool(info.done[0]) == doc["done"][i], f"done mismatch at step {i + 1}"          # :106
        assert bool(info.truncated[0]) == doc["truncated"][i], f"truncated mismatch at {i + 1}"  # :114
        assert np.max(np.abs(obs - np.asarray(doc["observation"][i]))) < 2e-7        # :124-133 (1e-4 there)
    env.close()


@pytest.mark.parametrize("name,kind", [("cartpole", 0), ("mountain_car", 1)])
def test_gym_against_python_batched(gym, oracle, name, kind):
    """The same 100 teacher-forced transitions as ONE batched step: env i replays fixture step i."""
    doc = load_golden(name)
    n, sd = 100, STATE_DIM[kind]
    start = np.zeros((sd, n), dtype=np.float32)
    for i in range(1, n):
        if not doc["done"][i - 1]:
            start[:, i] = doc["observation"][i - 1]
    env = gym.GpuVecEnv(kind, n, auto_reset=False)
    env.reset()
    env.set_state(dev(start))
    info = env.step(dev(np.asarray(doc["actions"], dtype=np.uint8)))
    obs = host(info.state)
    assert np.max(np.abs(obs.T - np.asarray(doc["observation"]))) < 2e-7
    assert host(info.done).tolist() == doc["done"]
    assert host(info.truncated).tolist() == doc["truncated"]
    assert host(info.reward).tolist() == doc["reward"]
    # and bit-exact against the oracle on the same inputs
    ref = oracle.VecState(kind, n, auto_reset=0)
    ref.reset()
    ref.state[:] = start
    o, r, f = ref.step(np.asarray(doc["actions"], dtype=np.uint8))
    assert_bit_equal(obs, o, "obs")
    env.close()


# ---------------------------------------------------------------------------------------------
# BASELINE config 1: one CartPole env, 10^4 random-action steps, injected reset states
# ---------------------------------------------------------------------------------------------
def test_cartpole_single_env_trace_1e4(gym, oracle):
    T = 10_000
    rng = np.random.default_rng(123)
    pool = rng.uniform(-0.05, 0.05, size=(4, T + 1)).astype(np.float32)
    actions = rng.integers(0, 2, size=T, dtype=np.uint8)
    env = gym.GpuVecEnv(0, 1, auto_reset=True, seed=1)
    env.set_reset_pool(dev(pool))
    env.set_state(dev(pool[:, :1]))
    ref = oracle.VecState(0, 1, auto_reset=1, seed=1)
    ref.set_reset_pool(pool)
    ref.state[:] = pool[:, :1]
    ref.sbt[:] = 0
    # per-call steps on the GPU ...
    a_dev = dev(actions.reshape(T, 1))
    got_obs = np.zeros((T, 4), np.float32)
    got_rew = np.zeros(T, np.float32)
    got_flg = np.zeros(T, np.uint8)
    got_steps = np.zeros(T, np.uint32)
    obs_all = torch.empty((T, 4, 1), device="cuda")
    rew_all = torch.empty((T, 1), device="cuda")
    flg_all = torch.empty((T, 1), dtype=torch.uint8, device="cuda")
    stp_all = torch.empty((T, 1), dtype=torch.int32, device="cuda")
    import ctypes as C
    lib = gym.load_library()
    for t in range(T):
        env.step_raw(a_dev[t], obs_all[t], rew_all[t], flg_all[t])
        lib.mgym_get_state(env._h, None, C.c_void_p(stp_all[t].data_ptr()), None, None)
    got_obs[:] = host(obs_all)[:, :, 0]
    got_rew[:] = host(rew_all)[:, 0]
    got_flg[:] = host(flg_all)[:, 0]
    got_steps[:] = host(stp_all)[:, 0].view(np.uint32)
    # ... against the oracle
    n_done = 0
    want_obs, want_rew = np.zeros_like(got_obs), np.zeros_like(got_rew)
    for t in range(T):
        o, r, f = ref.step(actions[t:t + 1])
        assert got_flg[t] == f[0], f"flags differ at step {t}"
        assert got_steps[t] == ref.steps[0], f"step counter differs at step {t}"
        assert_bit_equal(got_obs[t], o[:, 0], f"obs at step {t}")
        assert got_rew[t] == r[0]
        want_obs[t], want_rew[t] = o[:, 0], r[0]
        n_done += int(f[0] != 0)
    assert n_done > 300  # random policy: ~22 steps per episode
    # north_star's stated bar (1e-6 relative over the 10^4-step trace), on top of the bit equality above
    rel = np.max(np.abs(got_obs - want_obs) / np.maximum(np.abs(want_obs), 1e-30))
    assert rel <= REL_TOL[0]
    assert np.max(np.abs(got_rew - want_rew)) <= REL_TOL[0]
    s = env.stats()
    assert s.episodes == n_done == ref.stats.episodes and s.length_sum == ref.stats.length_sum
    env.close()


# ---------------------------------------------------------------------------------------------
# batched auto-reset parity, all five kinds, vector (N % 4 == 0) and scalar-lane (ragged N) paths
# ---------------------------------------------------------------------------------------------
def run_auto_parity(gym, oracle, kind, n, T, seed, use_pool=False, want_final=False, **cfg):
    rng = np.random.default_rng(1000 + kind)
    env = gym.GpuVecEnv(kind, n, auto_reset=True, seed=seed, track_returns=True, **cfg)
    ocfg = {k: v for k, v in cfg.items() if k in ("max_episode_steps", "sutton_barto_reward", "is_euler",
                                                    "goal_velocity", "env_index_base")}
    ocfg = {k: (int(v) if k != "goal_velocity" else v) for k, v in ocfg.items()}
    ref = oracle.VecState(kind, n, auto_reset=1, seed=seed, **ocfg)
    if use_pool:
        pool = random_states(rng, kind, 257)
        env.set_reset_pool(dev(pool))
        ref.set_reset_pool(pool)
    o0 = host(env.reset())
    assert_bit_equal(o0, ref.reset(), "reset obs")
    if kind != 0:  # start some envs from broad states too
        st = random_states(rng, kind, n)
        env.set_state(dev(st))
        ref.state[:] = st
        ref.steps[:] = 0
    for t in range(T):
        a = random_actions(rng, kind, n)
        if want_final:
            info, fin = env.step(dev(a), want_final_obs=True)
            o, r, f, fo = ref.step(a, want_final_obs=True)
            assert_bit_equal(host(fin), fo, f"final obs t={t}")
        else:
            info = env.step(dev(a))
            o, r, f = ref.step(a)
        flags = host(info.done).astype(np.uint8) | (host(info.truncated).astype(np.uint8) << 1)
        assert_bit_equal(flags, f, f"{KIND_NAMES[kind]} flags t={t}")
        assert_bit_equal(host(info.state), o, f"{KIND_NAMES[kind]} obs t={t}")
        assert_bit_equal(host(info.reward), r, f"{KIND_NAMES[kind]} reward t={t}")
    state, steps, _ = env.get_state()
    assert_bit_equal(host(state), ref.state, "final state")
    assert_bit_equal(host(steps).view(np.uint32), ref.steps, "step counters")
    s = env.stats()
    assert (s.episodes, s.terminated, s.truncated, s.length_sum) == (
        ref.stats.episodes, ref.stats.terminated, ref.stats.truncated, ref.stats.length_sum)
    assert s.return_sum == pytest.approx(ref.stats.return_sum, rel=1e-9, abs=1e-6)
    env.close()
    return s


@pytest.mark.parametrize("kind", range(5))
@pytest.mark.parametrize("n", [2048, 1027, 3076])  # whole TMA tiles; ragged scalar lanes; 3 tiles + a 4-env tail
def test_auto_reset_parity(gym, oracle, kind, n):
    T = {0: 600, 1: 300, 2: 300, 3: 250, 4: 200}[kind]
    s = run_auto_parity(gym, oracle, kind, n, T, seed=0x5EED)
    if kind in (0, 3, 4):
        assert s.episodes > 0


def test_cartpole_truncation_at_500(gym, oracle):
    """cartpole.rs:296-306: an env that survives 500 steps truncates with done=false, reward 1."""
    n = 64
    env = gym.GpuVecEnv(0, n, auto_reset=True, seed=5)
    ref = oracle.VecState(0, n, auto_reset=1, seed=5)
    env.reset(), ref.reset()
    z = np.zeros((4, n), np.float32)
    steps = np.full(n, 497, np.uint32)
    env.set_state(dev(z), dev(steps.view(np.int32)))
    ref.state[:] = z
    ref.steps[:] = steps
    seen_trunc = False
    for t in range(6):
        a = np.full(n, t % 2, np.uint8)
        info = env.step(dev(a))
        o, r, f = ref.step(a)
        flags = host(info.done).astype(np.uint8) | (host(info.truncated).astype(np.uint8) << 1)
        assert_bit_equal(flags, f, f"flags t={t}")
        assert_bit_equal(host(info.state), o, f"obs t={t}")
        if t == 2:
            assert (flags == 2).all() and (host(info.reward) == 1.0).all()
            seen_trunc = True
    assert seen_trunc


@pytest.mark.parametrize("kind,cfg", [
    (0, dict(sutton_barto_reward=True)),
    (0, dict(is_euler=False)),
    (1, dict(goal_velocity=0.02)),
    (1, dict(max_episode_steps=200)),
    (2, dict(max_episode_steps=50)),
    (0, dict(env_index_base=4096)),
    (4, dict(env_index_base=6)),       # unaligned slice base: scalar-lane path
])
def test_auto_reset_parity_options(gym, oracle, kind, cfg):
    run_auto_parity(gym, oracle, kind, 1024, 260, seed=99, **cfg)


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_auto_reset_injected_pool_and_final_obs(gym, oracle, kind):
    run_auto_parity(gym, oracle, kind, 1024, 200, seed=3, use_pool=True, want_final=True)


# ---------------------------------------------------------------------------------------------
# manual mode: the reference's own semantics, including stepping past termination
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", range(5))
@pytest.mark.parametrize("sb", [False, True])
@pytest.mark.parametrize("n", [1024, 3076, 515])  # one TMA tile; three tiles + a 4-env tail; scalar lanes
def test_manual_mode_parity(gym, oracle, kind, sb, n):
    if sb and kind != 0:
        pytest.skip("sutton_barto_reward is a CartPole option")
    T = 120
    rng = np.random.default_rng(50 + kind)
    env = gym.GpuVecEnv(kind, n, auto_reset=False, seed=11, sutton_barto_reward=sb)
    ref = oracle.VecState(kind, n, auto_reset=0, seed=11, sutton_barto_reward=int(sb))
    # before any reset the reference env holds a zero state and steps_beyond_terminated = Some(0)
    _, steps, sbt = env.get_state()
    assert_bit_equal(host(sbt).view(np.uint32), ref.sbt, "initial sbt")
    for t in range(T):
        if t == 40:   # caller-side reset of the envs that are done, cartpole.rs:468-470
            mask = (last_flags != 0).astype(np.uint8)
            assert_bit_equal(host(env.reset(mask=dev(mask))), ref.reset(mask=mask), "masked reset obs")
        if t == 80:
            assert_bit_equal(host(env.reset()), ref.reset(), "reset obs")
        a = random_actions(rng, kind, n)
        info = env.step(dev(a))
        o, r, f = ref.step(a)
        last_flags = host(info.done).astype(np.uint8) | (host(info.truncated).astype(np.uint8) << 1)
        assert_bit_equal(last_flags, f, f"flags t={t}")
        assert_bit_equal(host(info.state), o, f"obs t={t}")
        assert_bit_equal(host(info.reward), r, f"reward t={t}")
    state, steps, sbt = env.get_state()
    assert_bit_equal(host(steps).view(np.uint32), ref.steps, "steps_since_reset")
    if kind == 0:
        assert_bit_equal(host(sbt).view(np.uint32), ref.sbt, "steps_beyond_terminated")
        assert (ref.sbt > 1).any()  # the post-termination branch (cartpole.rs:330-347) was exercised
    env.close()


def test_mountain_car_wall_and_goal(gym, oracle):
    """mountain_car.rs:311-313 (left wall) and :318 (goal), which no reference fixture reaches."""
    st = np.array([[-1.2, -1.2, -1.19, 0.49, 0.5, 0.499],
                   [-0.01, 0.0, -0.07, 0.07, 0.0, 0.069]], dtype=np.float32)
    a = np.array([0, 2, 0, 2, 2, 2], dtype=np.uint8)
    env = gym.GpuVecEnv(1, 6, auto_reset=False)
    env.set_state(dev(st))
    info = env.step(dev(a))
    ref = oracle.VecState(1, 6, auto_reset=0)
    ref.state[:] = st
    o, r, f = ref.step(a)
    obs = host(info.state)
    assert_bit_equal(obs, o, "obs")
    assert obs[0, 0] == np.float32(-1.2) and obs[1, 0] == 0.0     # stopped by the wall
    assert host(info.done).tolist() == [bool(x & 1) for x in f]
    assert host(info.done)[3] and host(info.done)[4]               # reached the goal
    assert not host(info.truncated).any() and (host(info.reward) == -1.0).all()


# ---------------------------------------------------------------------------------------------
# rollout mode
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", range(5))
@pytest.mark.parametrize("n,policy", [(1024, False), (1024, True), (515, True), (515, False)])
def test_rollout_parity(gym, oracle, kind, n, policy):
    K = {0: 300, 1: 200, 2: 200, 3: 210, 4: 120}[kind]
    rng = np.random.default_rng(70 + kind)
    env = gym.GpuVecEnv(kind, n, auto_reset=True, seed=0xABCDEF, track_returns=True)
    ref = oracle.VecState(kind, n, auto_reset=1, seed=0xABCDEF)
    env.reset(), ref.reset()
    for chunk in range(2):  # two consecutive rollouts: the step index carries over
        a = None if policy else random_actions(rng, kind, (K, n))
        out = env.rollout(K, None if policy else dev(a))
        o, r, f, dc = ref.rollout(K, a)
        assert_bit_equal(host(out.flags), f, f"flags chunk {chunk}")
        assert_bit_equal(host(out.obs), o, f"obs chunk {chunk}")
        assert_bit_equal(host(out.reward), r, f"reward chunk {chunk}")
        assert int(out.done_count.item()) == dc == int((f != 0).sum())
    state, steps, _ = env.get_state()
    assert_bit_equal(host(state), ref.state, "state after rollout")
    assert_bit_equal(host(steps).view(np.uint32), ref.steps, "steps after rollout")
    s = env.stats()
    assert (s.episodes, s.length_sum) == (ref.stats.episodes, ref.stats.length_sum)
    assert s.return_sum == pytest.approx(ref.stats.return_sum, rel=1e-9, abs=1e-6)
    env.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_rollout_equals_repeated_steps(gym, kind):
    """Mode equivalence: K fused steps == K per-call steps with the sampled actions (bitwise)."""
    n, K = 4096, 64
    a_env = gym.GpuVecEnv(kind, n, seed=21)
    b_env = gym.GpuVecEnv(kind, n, seed=21)
    a_env.reset(), b_env.reset()
    out = a_env.rollout(K)
    for k in range(K):
        acts = b_env.sample_actions()
        info = b_env.step(acts)
        assert torch.equal(info.state.view(torch.int32), out.obs[k].view(torch.int32))
        assert torch.equal(info.reward, out.reward[k])
        flags = info.done.to(torch.uint8) | (info.truncated.to(torch.uint8) << 1)
        assert torch.equal(flags, out.flags[k])
    assert a_env.stats() == b_env.stats()


def test_sample_actions_matches_oracle(gym, oracle):
    for kind in range(5):
        env = gym.GpuVecEnv(kind, 64, seed=77, env_index_base=8)
        got = host(env.sample_actions())
        want = np.array([oracle.sample_action(kind, 77, 8 + i, 0) for i in range(64)], dtype=got.dtype)
        assert_bit_equal(got, want, KIND_NAMES[kind])
        if not CONTINUOUS[kind]:
            assert got.max() < NUM_ACTIONS[kind]


# ---------------------------------------------------------------------------------------------
# API behaviour
# ---------------------------------------------------------------------------------------------
def test_invalid_action_is_reported(gym):
    """cartpole.rs:392-403 / mountain_car.rs:374-385: the reference panics; the ABI returns an error."""
    for kind, bad in [(0, 2), (1, 3)]:
        env = gym.GpuVecEnv(kind, 8, validate_actions=True)
        env.reset()
        a = torch.zeros(8, dtype=torch.uint8, device="cuda")
        env.step(a)
        a[5] = bad
        before, t_before = [x.clone() for x in env.get_state()], env.step_index
        with pytest.raises(gym.InvalidActionError):
            env.step(a)
        with pytest.raises(gym.InvalidActionError):
            env.rollout(3, a.repeat(3, 1))
        with pytest.raises(gym.InvalidActionError):
            env.step_host(a.cpu(), None, None, None)
        # the reference asserts before any mutation (cartpole.rs:252): nothing was stepped
        assert env.step_index == t_before
        for x, y in zip(before, env.get_state()):
            assert torch.equal(x.view(torch.int32), y.view(torch.int32))
        env.step(torch.zeros(8, dtype=torch.uint8, device="cuda"))  # the handle stays usable
    env = gym.GpuVecEnv(0, 8)
    with pytest.raises(TypeError):
        env.step(torch.zeros(8, dtype=torch.int64, device="cuda"))   # wrong dtype (u32 scalar in the reference)
    with pytest.raises(ValueError):
        env.step(torch.zeros((8, 1), dtype=torch.uint8, device="cuda"))  # wrong rank, cartpole.rs:392-403


def test_stream_capture_is_refused(gym):
    """A captured step would bake one step index and one ticket range into the graph; the ABI refuses instead."""
    env = gym.GpuVecEnv(0, 4096, seed=1)
    env.reset()
    acts = env.sample_actions()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        graph.capture_begin()
        try:
            with pytest.raises(Exception, match="captured into a CUDA graph"):
                env.step(acts)
        finally:
            graph.capture_end()
    torch.cuda.synchronize()
    info = env.step(acts)   # the handle is still usable afterwards
    assert info.reward.shape == (4096,)
    env.close()


def test_reset_distribution_and_determinism(gym):
    """cartpole.rs:240: U[-0.05, 0.05); mountain_car.rs:281-283: U[-0.6,-0.4), v = 0; same seed => same stream
    (cartpole.rs:474-517), independent of how envs are split over handles."""
    n = 1 << 16
    e1 = gym.GpuVecEnv(0, n, seed=42)
    o1 = host(e1.reset()).copy()
    assert o1.min() >= -0.05 and o1.max() <= 0.05
    assert abs(o1.mean()) < 1e-3 and abs(o1.std() - 0.1 / np.sqrt(12)) < 1e-3
    e2 = gym.GpuVecEnv(0, n, seed=42)
    assert_bit_equal(host(e2.reset()), o1, "same seed")
    halves = [gym.GpuVecEnv(0, n // 2, seed=42, env_index_base=b) for b in (0, n // 2)]
    assert_bit_equal(np.concatenate([host(h.reset()) for h in halves], axis=1), o1, "sharded reset")
    e3 = gym.GpuVecEnv(0, n, seed=43)
    assert (host(e3.reset()) != o1).mean() > 0.99
    m = gym.GpuVecEnv(1, n, seed=42)
    om = host(m.reset())
    assert om[0].min() >= -0.6 and om[0].max() <= -0.4 and (om[1] == 0).all()


def test_sharded_handles_equal_one_handle(gym):
    """results are independent of the number of GPUs the env range is split over."""
    n, K = 8192, 50
    whole = gym.GpuVecEnv(0, n, seed=9)
    whole.reset()
    ow = whole.rollout(K)
    parts = [gym.GpuVecEnv(0, n // 4, seed=9, env_index_base=b * (n // 4)) for b in range(4)]
    outs = []
    for p in parts:
        p.reset()
        outs.append(p.rollout(K))
    assert torch.equal(torch.cat([o.obs for o in outs], dim=2).view(torch.int32), ow.obs.view(torch.int32))
    assert torch.equal(torch.cat([o.flags for o in outs], dim=1), ow.flags)
    tot = [sum(x) for x in zip(*[p.stats() for p in parts])]
    assert tuple(tot[:4]) == tuple(whole.stats()[:4])


@pytest.mark.parametrize("kind", range(5))
def test_scalar_adapter_protocol(gym, oracle, kind):
    """What bindings/rust/src/scalar.rs does, through the same entry points: ONE env in manual mode, host buffers only
    (mgym_reset + mgym_get_obs with a host pointer, mgym_step_host), the caller resetting after done or truncated
    exactly like the reference's own loops (cartpole.rs:460-471) -- compared with the oracle's scalar env."""
    env = gym.GpuVecEnv(kind, 1, auto_reset=False, seed=17, validate_actions=True)
    ref = oracle.VecState(kind, 1, auto_reset=0, seed=17)
    rng = np.random.default_rng(kind)
    obs, rew, flg = np.empty((env.obs_dim, 1), np.float32), np.empty(1, np.float32), np.empty(1, np.uint8)
    env.reset()
    assert_bit_equal(env.get_obs_host(), ref.reset(), "reset obs")
    episodes = 0
    for t in range(1100):  # beyond MountainCarContinuous' 999-step limit
        a = random_actions(rng, kind, 1)
        env.step_host(a, obs, rew, flg)
        o, r, f = ref.step(a)
        assert_bit_equal(obs, o, f"obs t={t}")
        assert_bit_equal(rew, r, f"reward t={t}")
        assert_bit_equal(flg, f, f"flags t={t}")
        if f[0]:
            episodes += 1
            env.reset()
            assert_bit_equal(env.get_obs_host(), ref.reset(), f"reset obs after t={t}")
    assert episodes > 0 or kind == 1  # MountainCar-v0 never truncates (mountain_car.rs:328)
    env.close()


def test_step_host_round_trip(gym, oracle):
    n = 4096
    env = gym.GpuVecEnv(0, n, seed=2)
    ref = oracle.VecState(0, n, auto_reset=1, seed=2)
    env.reset(), ref.reset()
    rng = np.random.default_rng(0)
    obs = torch.empty((4, n)).pin_memory()
    rew = torch.empty(n).pin_memory()
    flg = torch.empty(n, dtype=torch.uint8).pin_memory()
    for t in range(30):
        a = random_actions(rng, 0, n)
        env.step_host(torch.from_numpy(a).pin_memory(), obs, rew, flg)
        o, r, f = ref.step(a)
        assert_bit_equal(obs.numpy(), o, "obs")
        assert_bit_equal(flg.numpy(), f, "flags")
        assert_bit_equal(rew.numpy(), r, "reward")


@pytest.mark.parametrize("kind", range(5))
def test_checkpoint_resume_is_bit_identical(gym, kind):
    """SURVEY 5 (checkpoint / resume): save after K steps, restore into a fresh handle, continue: same bits as the
    uninterrupted run, including the Philox stream position and the statistics."""
    n, K = 4096, {0: 120, 1: 120, 2: 120, 3: 230, 4: 40}[kind]
    a = gym.GpuVecEnv(kind, n, seed=31)
    a.reset()
    a.rollout(K)
    blob = a.checkpoint()
    want = a.rollout(K)
    b = gym.GpuVecEnv(kind, n, seed=999)      # different seed on purpose: the checkpoint carries it
    b.restore(blob)
    got = b.rollout(K)
    assert torch.equal(got.obs.view(torch.int32), want.obs.view(torch.int32))
    assert torch.equal(got.flags, want.flags) and torch.equal(got.reward.view(torch.int32), want.reward.view(torch.int32))
    assert a.stats() == b.stats() and a.step_index == b.step_index
    with pytest.raises(gym.MgymError):
        gym.GpuVecEnv(kind, n // 2).restore(blob)


# ---------------------------------------------------------------------------------------------
# states outside every fast-path precondition: the reference-form fallback must give the oracle's bits
# ---------------------------------------------------------------------------------------------
def assert_equal_or_both_nan(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    both_nan = np.isnan(got) & np.isnan(want)
    same = (got.view(np.uint32) == want.view(np.uint32)) | both_nan
    if not same.all():
        idx = tuple(np.argwhere(~same)[0])
        raise AssertionError(f"{what}: {int((~same).sum())} differ; first at {idx}: got {got[idx]!r} want {want[idx]!r}")


def adversarial_states(kind, n, rng):
    sd = STATE_DIM[kind]
    big = np.array([0.0, -0.0, 1e-45, 1e-30, 0.7499, 0.75, 0.7501, 0.7853982, 1.5, 3.1415927, 10.0, 119.9, 120.0, 121.0,
                    1e4, 4194304.0, 1e7, 1e10, 1e20, 3e38, np.inf, np.nan], dtype=np.float32)
    s = np.zeros((sd, n), dtype=np.float32)
    for c in range(sd):
        s[c] = rng.choice(big, size=n) * rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=n)
    # keep a fraction ordinary so that fast and fallback envs share warps and threads
    ordinary = rng.random(n) < 0.5
    s[:, ordinary] = random_states(rng, kind, int(ordinary.sum()))
    return s


@pytest.mark.parametrize("kind", range(5))
@pytest.mark.parametrize("auto", [True, False])
def test_fallback_paths_match_oracle(gym, oracle, kind, auto):
    n = 4096
    rng = np.random.default_rng(900 + kind)
    env = gym.GpuVecEnv(kind, n, auto_reset=auto, seed=4)
    ref = oracle.VecState(kind, n, auto_reset=int(auto), seed=4)
    env.reset(), ref.reset()
    with np.errstate(all="ignore"):
        for rep in range(3):
            st = adversarial_states(kind, n, rng)
            env.set_state(dev(st))
            ref.state[:] = st
            ref.steps[:] = 0
            ref.sbt[:] = 0
            ref.ep_return[:] = 0
            for t in range(3):
                a = random_actions(rng, kind, n)
                info = env.step(dev(a))
                o, r, f = ref.step(a)
                flags = host(info.done).astype(np.uint8) | (host(info.truncated).astype(np.uint8) << 1)
                assert_bit_equal(flags, f, f"{KIND_NAMES[kind]} flags rep {rep} t {t}")
                assert_equal_or_both_nan(host(info.state), o, f"{KIND_NAMES[kind]} obs rep {rep} t {t}")
                assert_equal_or_both_nan(host(info.reward), r, f"{KIND_NAMES[kind]} reward rep {rep} t {t}")
    env.close()


@pytest.mark.parametrize("kind", range(5))
@pytest.mark.parametrize("n", [2048, 515])
def test_rollout_adversarial_states_match_oracle(gym, oracle, kind, n):
    """The fused rollout from NaN / inf / huge injected states: its reference-form fallbacks, and for MountainCar the
    entry test that decides between the trusted and the checked loop, give the oracle's results."""
    K = 6
    rng = np.random.default_rng(1700 + kind)
    env = gym.GpuVecEnv(kind, n, auto_reset=True, seed=6)
    ref = oracle.VecState(kind, n, auto_reset=1, seed=6)
    env.reset(), ref.reset()
    with np.errstate(all="ignore"):
        for rep in range(2):
            st = adversarial_states(kind, n, rng)
            if rep == 1:  # whole warps of ordinary states next to whole warps of adversarial ones
                st[:, : n // 2] = random_states(rng, kind, n // 2)
            env.set_state(dev(st))
            ref.state[:] = st
            ref.steps[:] = 0
            ref.ep_return[:] = 0
            a = random_actions(rng, kind, (K, n))
            out = env.rollout(K, dev(a))
            o, r, f, dc = ref.rollout(K, a)
            assert_bit_equal(host(out.flags), f, f"{KIND_NAMES[kind]} flags rep {rep}")
            assert_equal_or_both_nan(host(out.obs), o, f"{KIND_NAMES[kind]} obs rep {rep}")
            assert_equal_or_both_nan(host(out.reward), r, f"{KIND_NAMES[kind]} reward rep {rep}")
            assert int(out.done_count.item()) == dc
    env.close()


def test_mountain_car_rollout_leaves_trusted_loop_on_bad_pool_state(gym, oracle):
    """MountainCar's rollout runs without per-step precondition tests once its entry invariant holds; a reset
    state injected through the pool can break the invariant, and the kernel must notice (it re-tests after pool
    resets) and continue in the checked loop.  Short episodes (TimeLimit 3) make every env reset from a pool
    that holds NaN, inf and huge positions next to ordinary ones."""
    n, K = 4096, 24
    rng = np.random.default_rng(77)
    env = gym.GpuVecEnv(1, n, auto_reset=True, seed=9, max_episode_steps=3)
    ref = oracle.VecState(1, n, auto_reset=1, seed=9, max_episode_steps=3)
    pool = random_states(rng, 1, 257)
    bad = adversarial_states(1, 257, rng)
    pick = rng.random(257) < 0.1
    pool[:, pick] = bad[:, pick]
    env.set_reset_pool(dev(pool))
    ref.set_reset_pool(pool)
    env.reset(), ref.reset()
    st = random_states(rng, 1, n)      # ordinary: every warp enters the trusted loop
    env.set_state(dev(st))
    ref.state[:] = st
    ref.steps[:] = 0
    with np.errstate(all="ignore"):
        for chunk in range(2):
            a = random_actions(rng, 1, (K, n))
            out = env.rollout(K, dev(a))
            o, r, f, dc = ref.rollout(K, a)
            assert_bit_equal(host(out.flags), f, f"flags chunk {chunk}")
            assert_equal_or_both_nan(host(out.obs), o, f"obs chunk {chunk}")
            assert int(out.done_count.item()) == dc
    s = env.stats()
    assert (s.episodes, s.truncated, s.length_sum) == (ref.stats.episodes, ref.stats.truncated, ref.stats.length_sum)
    env.close()


def test_mountain_car_continuous_rollout_survives_nan_and_inf_actions(gym, oracle):
    """MountainCarContinuous runs the trusted loop too, which additionally needs actions that are not NaN: a warp
    that meets one runs that step checked and re-tests its invariant.  NaN, +-inf and huge actions are sprinkled
    over ordinary ones; flags, observations and rewards must match the oracle (NaN where it has NaN)."""
    n, K = 4096, 20
    rng = np.random.default_rng(4242)
    env = gym.GpuVecEnv(2, n, auto_reset=True, seed=13)
    ref = oracle.VecState(2, n, auto_reset=1, seed=13)
    env.reset(), ref.reset()
    with np.errstate(all="ignore"):
        for chunk in range(3):
            a = random_actions(rng, 2, (K, n))
            weird = np.array([np.nan, np.inf, -np.inf, 3e38, -3e38, 0.0, -0.0], dtype=np.float32)
            hit = rng.random((K, n)) < (0.0 if chunk == 0 else 0.002)   # first chunk: pure trusted loop
            a[hit] = rng.choice(weird, size=int(hit.sum()))
            out = env.rollout(K, dev(a))
            o, r, f, dc = ref.rollout(K, a)
            assert_bit_equal(host(out.flags), f, f"flags chunk {chunk}")
            assert_equal_or_both_nan(host(out.obs), o, f"obs chunk {chunk}")
            assert_equal_or_both_nan(host(out.reward), r, f"reward chunk {chunk}")
            assert int(out.done_count.item()) == dc
    env.close()


def test_mountain_car_1000_step_chunked_rollout(gym, oracle):
    """BASELINE configs[2]: a 1000-step MountainCar rollout run as 32-step launches over one reused ring equals
    the oracle's single 1000-step rollout (flags, rewards, observations of every step, final state)."""
    n, T = 512, 1000
    env = gym.GpuVecEnv(1, n, seed=8)
    ref = oracle.VecState(1, n, auto_reset=1, seed=8)
    env.reset(), ref.reset()
    o, r, f, dones = ref.rollout(T)
    total_done = 0
    for first, out in env.iter_rollout(T, chunk=32):
        k = out.flags.shape[0]
        assert_bit_equal(host(out.obs), o[first:first + k], f"obs of steps [{first},{first + k})")
        assert_bit_equal(host(out.flags), f[first:first + k], "flags")
        assert_bit_equal(host(out.reward), r[first:first + k], "reward")
        total_done += int(out.done_count.item())
    assert total_done == dones and env.step_index == T
    state, steps, _ = env.get_state()
    assert_bit_equal(host(state), ref.state, "final state")


@pytest.mark.parametrize("kind", [0, 3])
def test_step_host_pipelined_chunks_equal_plain_step(gym, kind):
    """mgym_step_host cuts large batches into chunks on two internal streams (H2D / kernel / D2H overlap); the
    result must equal the ordinary device-buffer step bit for bit, including a ragged last chunk."""
    n = (1 << 20) + 3 * 1024 + 8
    a_env, b_env = gym.GpuVecEnv(kind, n, seed=12), gym.GpuVecEnv(kind, n, seed=12)
    a_env.reset(), b_env.reset()
    od = a_env.obs_dim
    obs = torch.empty((od, n)).pin_memory()
    rew = torch.empty(n).pin_memory()
    flg = torch.empty(n, dtype=torch.uint8).pin_memory()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    for t in range(25):
        if a_env.continuous:
            acts = torch.rand(n, device="cuda", generator=gen) * 4 - 2
        else:
            acts = torch.randint(0, 2, (n,), dtype=torch.uint8, device="cuda", generator=gen)
        a_env.step_host(acts.cpu().pin_memory(), obs, rew, flg)
        info = b_env.step(acts)
        assert torch.equal(obs.cuda().view(torch.int32), info.state.view(torch.int32)), f"obs differ at step {t}"
        assert torch.equal(rew.cuda().view(torch.int32), info.reward.view(torch.int32))
        assert torch.equal(flg.cuda(), info.flags)
    assert a_env.stats() == b_env.stats() and a_env.step_index == b_env.step_index
