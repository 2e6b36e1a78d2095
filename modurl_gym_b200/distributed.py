"""Multi-GPU plumbing: envs shard as independent contiguous slices; the only
collective is a sum all-reduce of the 5-double episode-statistics vector.  Pure host logic, so it
runs on CPU tensors with gloo as well as on CUDA tensors with NCCL."""
from collections import namedtuple

import torch
import torch.distributed as dist

EpisodeStats = namedtuple("EpisodeStats", ["episodes", "terminated", "truncated", "length_sum", "return_sum"])


def shard_range(total_envs, rank, world_size, multiple=1024):
    """Global env range [begin, end) owned by `rank`: contiguous, sizes differ by at most `multiple`,
    every boundary a multiple of `multiple` (the TMA tile) so each slice keeps the vector path.
    The Philox streams are keyed by the GLOBAL env index, so results do not depend on world_size."""
    if total_envs % multiple:
        multiple = 4 if total_envs % 4 == 0 else 1
    units = total_envs // multiple
    base, extra = divmod(units, world_size)
    begin = (rank * base + min(rank, extra)) * multiple
    size = (base + (1 if rank < extra else 0)) * multiple
    return begin, begin + size


def all_reduce_stats_vector(vec, group=None):
    """Sum {episodes, terminated, truncated, length_sum, return_sum} over all ranks, in place."""
    assert vec.dtype == torch.float64 and vec.numel() == 5
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    v = vec.tolist()
    return EpisodeStats(int(v[0]), int(v[1]), int(v[2]), int(v[3]), v[4])


def max_over_ranks(value, device, group=None):
    """Device-timed durations are reported as the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
