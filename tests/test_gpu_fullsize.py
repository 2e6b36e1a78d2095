"""BASELINE.json's full sizes (2^24 envs; 2^22 for Acrobot/Pendulum), where the scalar oracle cannot replay
everything: size-independent properties plus exact oracle replays of slices of the big batch.

  * slices: envs [0,1024), a middle window and the last 1024 of the batch are replayed by the oracle with
    their GLOBAL env indices (the Philox streams are keyed by them) and must match bit for bit;
  * mode equivalence: K fused rollout steps == K per-call steps, bitwise, on the whole batch;
  * determinism: same seed => same bits; conservation: episode statistics == what the flags say;
  * domain invariants (bounds, rewards) on every element."""
import numpy as np
import pytest

from helpers import CONTINUOUS, KIND_NAMES, OBS_DIM, assert_bit_equal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

N_BIG = {0: 1 << 24, 1: 1 << 24, 2: 1 << 24, 3: 1 << 22, 4: 1 << 22}


@pytest.fixture(scope="module")
def gym():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import modurl_gym_b200 as m

    m.load_library()
    return m


def windows(n):
    return [(0, 1024), (n // 2 - 512 + 4, n // 2 + 512 + 4), (n - 1024, n)]


@pytest.mark.parametrize("kind", range(5))
def test_full_size_slices_match_oracle(gym, oracle, kind):
    n, K, seed = N_BIG[kind], {0: 64, 1: 64, 2: 64, 3: 48, 4: 24}[kind], 0xC0FFEE + kind
    env = gym.GpuVecEnv(kind, n, seed=seed)
    env.reset()
    # device policy: actions are a pure function of (seed, global env index, step), so the oracle can
    # regenerate them for any slice
    out = env.rollout(K)
    # then per-call steps with explicit actions sampled on the device
    acts = env.sample_actions()
    info = env.step(acts)
    torch.cuda.synchronize()
    for (a, b) in windows(n):
        ref = oracle.VecState(kind, b - a, auto_reset=1, seed=seed, env_index_base=a)
        ref.reset()
        o, r, f, dones = ref.rollout(K)
        assert_bit_equal(out.obs[:, :, a:b].cpu().numpy(), o, f"{KIND_NAMES[kind]} rollout obs [{a},{b})")
        assert_bit_equal(out.flags[:, a:b].cpu().numpy(), f, "rollout flags")
        assert_bit_equal(out.reward[:, a:b].cpu().numpy(), r, "rollout reward")
        want_a = np.array([oracle.sample_action(kind, seed, a + i, K) for i in range(b - a)],
                          dtype=np.float32 if CONTINUOUS[kind] else np.uint8)
        assert_bit_equal(acts[a:b].cpu().numpy(), want_a, "sampled actions")
        o1, r1, f1 = ref.step(want_a)
        assert_bit_equal(info.state[:, a:b].cpu().numpy(), o1, "step obs")
        flags = (info.done.to(torch.uint8) | (info.truncated.to(torch.uint8) << 1))[a:b].cpu().numpy()
        assert_bit_equal(flags, f1, "step flags")
    # conservation: statistics == what the flags say
    s = env.stats()
    n_done = int((out.flags != 0).sum().item()) + int((info.done | info.truncated).sum().item())
    assert s.episodes == n_done == int(out.done_count.item()) + int((info.done | info.truncated).sum().item())
    env.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_full_size_rollout_equals_steps(gym, kind):
    n, K = N_BIG[kind], 6
    a, b = gym.GpuVecEnv(kind, n, seed=5), gym.GpuVecEnv(kind, n, seed=5)
    a.reset(), b.reset()
    out = a.rollout(K)
    for k in range(K):
        info = b.step(b.sample_actions())
        assert torch.equal(info.state.view(torch.int32), out.obs[k].view(torch.int32)), f"obs differ at step {k}"
        assert torch.equal(info.reward, out.reward[k])
        assert torch.equal(info.done.to(torch.uint8) | (info.truncated.to(torch.uint8) << 1), out.flags[k])
    sa, sb = a.get_state(), b.get_state()
    assert torch.equal(sa[0].view(torch.int32), sb[0].view(torch.int32)) and torch.equal(sa[1], sb[1])
    assert a.stats() == b.stats()


def test_full_size_cartpole_invariants_and_determinism(gym):
    n, T = N_BIG[0], 40
    e1, e2 = gym.GpuVecEnv(0, n, seed=77), gym.GpuVecEnv(0, n, seed=77)
    e1.reset(), e2.reset()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    total_done = 0
    for t in range(T):
        acts = torch.randint(0, 2, (n,), dtype=torch.uint8, device="cuda", generator=gen)
        i1, f1 = e1.step(acts, want_final_obs=True)
        i2 = e2.step(acts)
        assert torch.equal(i1.state.view(torch.int32), i2.state.view(torch.int32))   # determinism
        done = i1.done | i1.truncated
        total_done += int(done.sum().item())
        assert bool((i1.reward == 1.0).all())                                        # cartpole.rs:310-329
        assert not bool((i1.done & i1.truncated).any())                              # :297-306 excludes both
        # terminal observations are outside a threshold, live ones inside (cartpole.rs:291-294)
        thr = np.array([0x3e567750], dtype=np.uint32).view(np.float32)[0]
        outside = (f1[0].abs() > 2.4) | (f1[2].abs() > float(thr))
        assert torch.equal(outside, i1.done)
        # a reset env restarts inside U[-0.05, 0.05) (cartpole.rs:240)
        assert bool((i1.state[:, done].abs() <= 0.05).all())
    s = e1.stats()
    assert s.episodes == total_done and s.truncated == 0 and s.terminated == total_done
    assert s.return_sum == float(s.length_sum)                                       # reward 1 per step
    _, steps, _ = e1.get_state()
    assert int(steps.max().item()) <= T


def test_full_size_mountain_car_invariants(gym):
    n, K = N_BIG[1], 100
    env = gym.GpuVecEnv(1, n, seed=3)
    env.reset()
    out = env.rollout(K, want_obs=True, obs=torch.empty((K, 2, n), device="cuda"))
    pos, vel = out.obs[:, 0], out.obs[:, 1]
    assert float(pos.min()) >= -1.2000000477 and float(pos.max()) <= 0.6000000239   # mountain_car.rs:308
    assert float(vel.abs().max()) <= 0.0700000003                                    # :304
    assert bool((out.reward == -1.0).all())                                          # :319
    assert not bool((out.flags & 2).any())                                           # :328 never truncates
    wall = (pos == np.float32(-1.2))
    assert bool((vel[wall] >= 0).all())                                              # :311-313


def test_two_to_the_27_envs_on_one_gpu(gym, oracle):
    """BASELINE configs[4]'s whole population (2^27 envs) on ONE handle: element offsets pass 2^31 and byte
    offsets 2^33, so any 32-bit index arithmetic in the kernels would show here.  The first and the last 2048 envs
    are replayed by the oracle (global indices key the Philox streams), for a fused rollout and for per-call steps."""
    n, K, w = 1 << 27, 6, 2048
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs 24 GB of free device memory")
    env = gym.GpuVecEnv(0, n, seed=77)
    env.reset()
    out = env.rollout(K)
    for a in (0, n - w):
        ref = oracle.VecState(0, w, auto_reset=1, seed=77, env_index_base=a)
        ref.reset()
        o, r, f, dc = ref.rollout(K)
        assert_bit_equal(out.obs[:, :, a:a + w].cpu().numpy(), o, f"rollout obs of envs [{a},{a + w})")
        assert_bit_equal(out.flags[:, a:a + w].cpu().numpy(), f, "rollout flags")
        for t in range(3):
            acts = env.sample_actions()
            info = env.step(acts)
            o1, r1, f1 = ref.step(acts[a:a + w].cpu().numpy())
            assert_bit_equal(info.state[:, a:a + w].cpu().numpy(), o1, f"step obs of envs [{a},{a + w}) t={t}")
            assert_bit_equal(info.flags[a:a + w].cpu().numpy(), f1, "step flags")
        # the next window starts from the same point in time: rewind is not possible, so use a fresh handle
        if a == 0:
            env.close()
            del out
            torch.cuda.empty_cache()
            env = gym.GpuVecEnv(0, n, seed=77)
            env.reset()
            out = env.rollout(K)
    env.close()
