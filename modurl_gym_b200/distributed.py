"""Multi-GPU plumbing: envs shard as independent contiguous slices; the only
collective is a sum all-reduce of the 5-double episode-statistics vector.  Pure host logic, so it
runs on CPU tensors with gloo as well as on CUDA tensors with NCCL."""
import ctypes as C
import os
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


# -------------------------------------------------------------------------------------------------
# A raw ncclComm_t for the C ABI's own collective (mgym_stats_allreduce), next to torch.distributed's.
# -------------------------------------------------------------------------------------------------
class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]  # nccl.h: NCCL_UNIQUE_ID_BYTES


def loaded_nccl_path():
    """Path of the libnccl this process already has mapped (torch's bundled one once a NCCL process group
    exists), so that the communicator created here and the ncclAllReduce that libmgym.so resolves with dlsym
    come from the SAME library; None if none is mapped."""
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                path = line.rsplit(" ", 1)[-1].strip()
                if os.path.basename(path).startswith("libnccl.so"):
                    return path
    except OSError:
        pass
    return None


def unique_id_bytes(lib):
    uid = _NcclUniqueId()
    rc = lib.ncclGetUniqueId(C.byref(uid))
    if rc != 0:
        raise RuntimeError(f"ncclGetUniqueId returned {rc}")
    return bytes(bytearray(uid.internal))


def exchange_unique_id(make_id, rank, device, group=None):
    """Rank 0 draws the 128-byte ncclUniqueId, every rank receives it through torch.distributed (any backend)."""
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


class NativeNcclComm:
    """ncclCommInitRank over the ranks of the default process group; `handle` is the ncclComm_t to pass to
    mgym_stats_allreduce.  One communicator per process; call close() before the process group goes away."""

    def __init__(self, rank, world_size, device, group=None):
        path = loaded_nccl_path() or "libnccl.so.2"
        # RTLD_GLOBAL: promotes the already-mapped library so that dlsym(RTLD_DEFAULT, "ncclAllReduce") inside
        # libmgym.so finds this very copy
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        self.path = path
        self.lib.ncclGetUniqueId.argtypes = [C.POINTER(_NcclUniqueId)]
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        raw = exchange_unique_id(lambda: unique_id_bytes(self.lib), rank, device, group)
        uid = _NcclUniqueId()
        C.memmove(C.byref(uid), raw, 128)
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            rc = self.lib.ncclCommInitRank(C.byref(self.handle), int(world_size), uid, int(rank))
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank returned {rc} ({path})")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ncclCommDestroy(self.handle)
            self.handle = None
