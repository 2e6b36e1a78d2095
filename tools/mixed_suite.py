#!/usr/bin/env python
"""BASELINE.json configs[4]: the mixed classic-control suite sharded over the GPUs of one box.

Every rank (one per GPU, torchrun) hosts an equal slice of every kind -- 2^24 envs per GPU in total
(3 x 2^22 CartPole / MountainCar / MountainCarContinuous + 2 x 2^21 Pendulum / Acrobot), 2^27 on 8 GPUs --
steps them with the device-side random policy (fused rollouts), and the per-kind episode statistics
{episodes, terminated, truncated, length_sum, return_sum} are summed over ranks with ONE NCCL all-reduce of a
5 x 5 double matrix.  Global env indices key the Philox streams, so the union of the slices is the same
population whatever the number of GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mixed_suite.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import modurl_gym_b200 as m  # noqa: E402
from modurl_gym_b200.distributed import max_over_ranks  # noqa: E402

NAMES = ["CartPole-v1", "MountainCar-v0", "MountainCarContinuous-v0", "Pendulum-v1", "Acrobot-v1"]
PER_GPU = [1 << 22, 1 << 22, 1 << 22, 1 << 21, 1 << 21]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every slice (smoke runs)")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    envs, bufs = [], []
    for kind, n in enumerate(PER_GPU):
        n = max(1024, int(n * args.scale) // 1024 * 1024)
        env = m.GpuVecEnv(kind, n, device=local, seed=0x5EED, env_index_base=rank * n, track_returns=True)
        env.reset()
        envs.append(env)
        K = args.chunk
        bufs.append((torch.empty((K, env.obs_dim, n), device=dev), torch.empty((K, n), device=dev),
                     torch.empty((K, n), dtype=torch.uint8, device=dev)))

    def sweep():
        for env, (o, r, f) in zip(envs, bufs):
            env.rollout(args.chunk, None, obs=o, reward=r, flags=f, count_done=False)

    sweep()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # per-kind timing (one kind at a time) and the whole sweep
    per_kind = []
    for env, (o, r, f) in zip(envs, bufs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.chunks):
            env.rollout(args.chunk, None, obs=o, reward=r, flags=f, count_done=False)
        e1.record()
        torch.cuda.synchronize()
        per_kind.append(max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.chunks):
        sweep()
    e1.record()
    torch.cuda.synchronize()
    total_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    # ONE collective: 5 kinds x 5 statistics
    mat = torch.stack([env.stats_tensor() for env in envs])
    if world > 1:
        dist.all_reduce(mat, op=dist.ReduceOp.SUM)
    if rank == 0:
        steps = args.chunk * args.chunks
        rows = {}
        for kind, env in enumerate(envs):
            ep, term, trunc, length, ret = mat[kind].tolist()
            rows[NAMES[kind]] = {
                "envs_total": env.num_envs * world, "env_steps_per_s": env.num_envs * world * steps / per_kind[kind],
                "episodes": int(ep), "terminated": int(term), "truncated": int(trunc),
                "mean_length": length / max(ep, 1), "mean_return": ret / max(ep, 1)}
        total_envs = sum(e.num_envs for e in envs) * world
        print(json.dumps({"config": "mixed classic_control suite (BASELINE configs[4])", "n_gpus": world,
                          "envs_total": total_envs, "steps_per_env": steps, "mode": "fused rollout, device policy",
                          "suite_env_steps_per_s": total_envs * steps / total_s, "per_kind": rows,
                          "collective": "one NCCL all-reduce (sum) of a 5x5 float64 matrix"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
