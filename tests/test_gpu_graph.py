"""CUDA-graph-capturable handles (cfg.device_clock = 1): the step index and the tile tickets live on the device,
so a captured [sample_actions, step] or rollout launch replays correctly.  Every comparison is against a default
handle run eagerly with the same seed: the two must agree bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gym():
    import modurl_gym_b200 as m

    m.load_library()
    return m


def same(a, b):
    if a.dtype == torch.float32:
        a, b = a.view(torch.int32), b.view(torch.int32)
    return torch.equal(a, b)


def flags_of(info):
    return info.flags  # not .done / .truncated: StepInfo caches those, and a replayed graph rewrites the buffer


@pytest.mark.parametrize("kind", [0, 1, 3, 4])
@pytest.mark.parametrize("n", [5120, 1027, 3076])  # TMA tiles; scalar lanes; tiles + a vector tail
def test_device_clock_handle_equals_default_handle_eagerly(gym, kind, n):
    a = gym.GpuVecEnv(kind, n, seed=11)
    b = gym.GpuVecEnv(kind, n, seed=11, graph_capturable=True)
    assert same(a.reset().clone(), b.reset().clone())
    for t in range(40):
        acts = a.sample_actions()
        assert same(acts, b.sample_actions()), f"sampled actions differ at step {t}"
        ia, ib = a.step(acts), b.step(acts)
        assert same(ia.state, ib.state) and same(ia.reward, ib.reward) and same(flags_of(ia), flags_of(ib)), t
    ra, rb = a.rollout(24), b.rollout(24)
    assert same(ra.obs, rb.obs) and same(ra.reward, rb.reward) and same(ra.flags, rb.flags)
    assert int(ra.done_count.item()) == int(rb.done_count.item())
    assert a.step_index == b.step_index == 64
    assert a.stats() == b.stats()
    a.close(), b.close()


@pytest.mark.parametrize("kind,n", [(0, 4096), (0, 1027), (1, 8192), (3, 2048)])
def test_graph_replay_of_sample_and_step_matches_eager(gym, kind, n):
    T = 50
    a = gym.GpuVecEnv(kind, n, seed=5)
    b = gym.GpuVecEnv(kind, n, seed=5, graph_capturable=True)
    a.reset(), b.reset()
    acts_b = torch.empty(n, dtype=b.action_dtype, device="cuda")
    # one eager step first (occupancy queries, buffer set-up), mirrored on the default handle
    b.step(b.sample_actions(out=acts_b))
    a.step(a.sample_actions())
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        b.sample_actions(out=acts_b)
        info_b = b.step(acts_b)
    # capture does not execute: the env is where the eager step left it
    for t in range(T):
        graph.replay()
        info_a = a.step(a.sample_actions())
        assert same(info_a.state, info_b.state), f"state differs at replay {t}"
        assert same(info_a.reward, info_b.reward) and same(flags_of(info_a), flags_of(info_b)), t
    assert b.step_index == a.step_index == T + 1
    assert a.stats() == b.stats()
    a.close(), b.close()


def test_graph_replay_of_rollout_matches_eager(gym):
    n, K, R = 4096, 16, 8
    a = gym.GpuVecEnv(0, n, seed=9)
    b = gym.GpuVecEnv(0, n, seed=9, graph_capturable=True)
    a.reset(), b.reset()
    obs = torch.empty((K, 4, n), device="cuda")
    rew = torch.empty((K, n), device="cuda")
    flg = torch.empty((K, n), dtype=torch.uint8, device="cuda")
    b.rollout(K, obs=obs, reward=rew, flags=flg)
    a.rollout(K)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        b.rollout(K, obs=obs, reward=rew, flags=flg, count_done=False)
    for r in range(R):
        graph.replay()
        out = a.rollout(K)
        assert same(out.obs, obs) and same(out.reward, rew) and same(out.flags, flg), f"replay {r}"
    assert b.step_index == a.step_index == (R + 1) * K
    a.close(), b.close()


def test_default_handle_still_refuses_capture_and_checkpoint_carries_the_clock(gym):
    n = 2048
    b = gym.GpuVecEnv(0, n, seed=3, graph_capturable=True)
    b.reset()
    for _ in range(7):
        b.step(b.sample_actions())
    blob = b.checkpoint()
    c = gym.GpuVecEnv(0, n, seed=3, graph_capturable=True)
    c.restore(blob)
    assert c.step_index == 7
    for _ in range(5):
        acts = b.sample_actions()
        assert same(acts, c.sample_actions())
        assert same(b.step(acts).state.clone(), c.step(acts).state)
    b.close(), c.close()


@pytest.mark.parametrize("n", [4096, 1 << 20])  # one launch; eight pipelined chunk launches over two streams
def test_step_host_with_a_device_clock(gym, n):
    """mgym_step_host on a capturable handle: every chunk launch reads the step index from the device clock, which
    one clock_advance_kernel moves on after the chunks have joined; results equal the default handle's."""
    import numpy as np

    a = gym.GpuVecEnv(0, n, seed=21)
    b = gym.GpuVecEnv(0, n, seed=21, graph_capturable=True)
    a.reset(), b.reset()
    rng = np.random.default_rng(5)
    bufs = []
    for _ in range(2):
        bufs.append((torch.empty((4, n)).pin_memory(), torch.empty(n).pin_memory(),
                     torch.empty(n, dtype=torch.uint8).pin_memory()))
    for t in range(30):
        acts = torch.from_numpy(rng.integers(0, 2, n, dtype=np.uint8)).pin_memory()
        a.step_host(acts, *bufs[0])
        b.step_host(acts, *bufs[1])
        for x, y in zip(bufs[0], bufs[1]):
            assert same(x, y), f"host outputs differ at step {t}"
        if t % 7 == 3:  # interleave device-buffer steps: both kinds of call advance the same clock
            d = acts.cuda()
            assert same(a.step(d).state, b.step(d).state)
    assert a.step_index == b.step_index
    assert a.stats() == b.stats()
    a.close(), b.close()
