"""World-size-2 gloo test of the multi-GPU host logic (no GPU needed): slices are independent, the
Philox streams are keyed by the global env index, and the only collective is the statistics all-reduce.
Each rank plays its slice with the CPU oracle standing in for the device (the oracle is the checker of
the sharding arithmetic here, not a product path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, STEPS, SEED = 4096, 120, 0xBEEF


def play(begin, end):
    from oracle import oracle as o

    vs = o.VecState(o.CARTPOLE, end - begin, auto_reset=1, seed=SEED, env_index_base=begin)
    vs.reset()
    out = vs.rollout(STEPS)  # device policy: actions keyed by the global env index
    s = vs.stats
    return out, np.array([s.episodes, s.terminated, s.truncated, s.length_sum, s.return_sum], dtype=np.float64)


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from modurl_gym_b200.distributed import all_reduce_stats_vector, exchange_unique_id, max_over_ranks, shard_range

    begin, end = shard_range(TOTAL, rank, world)
    (obs, rew, flg, dones), vec = play(begin, end)
    total = all_reduce_stats_vector(torch.from_numpy(vec.copy()))
    slowest = max_over_ranks(10.0 + rank, "cpu")
    # the rendezvous of the native NCCL communicator (NativeNcclComm): rank 0 draws the 128-byte id, every rank
    # receives it through the process group; only rank 0's generator may run
    def make_id():
        assert rank == 0
        return bytes((7 * i + 3) & 0xFF for i in range(128))

    uid = exchange_unique_id(make_id, rank, "cpu")
    q.put((rank, begin, end, obs[-1], int(dones), tuple(total), slowest, uid))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (whole_obs, _, _, whole_dones), whole_vec = play(0, TOTAL)
    assert [g[1:3] for g in got] == [(0, 2048), (2048, 4096)]
    # concatenated slices == the unsharded run, bit for bit
    cat = np.concatenate([g[3] for g in got], axis=1)
    assert np.array_equal(cat.view(np.uint32), whole_obs[-1].view(np.uint32))
    assert sum(g[4] for g in got) == whole_dones
    for g in got:  # every rank holds the same reduced statistics
        assert g[5][:4] == tuple(int(x) for x in whole_vec[:4])
        assert g[5][4] == pytest.approx(whole_vec[4])
        assert g[6] == 11.0
        assert g[7] == bytes((7 * i + 3) & 0xFF for i in range(128))  # both ranks hold rank 0's unique id


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from modurl_gym_b200.distributed import shard_range

    for total, world in [(1 << 27, 8), (1 << 24, 1), (1 << 24, 3), (4096, 2), (1000, 3), (7, 2)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        if total % 1024 == 0:
            assert all(b % 1024 == 0 for b, _ in spans)
    assert shard_range(1 << 27, 3, 8) == (3 << 24, 4 << 24)
