"""Host-side sharding of a vector of envs over the GPUs of one box.

Envs never read each other's state (env_binding.h:520-522 is a loop over disjoint Env*),
so the path shards by contiguous env-index ranges with NO per-step communication.  The only
exchange is vec_log: every rank's integer episode sums are all-reduced (NCCL on GPUs; any
torch.distributed backend works, the CPU tests use gloo) and then averaged exactly like
env_binding.h:572-591 does for one process.  Reset streams are keyed by the GLOBAL env id
(`env_id_base`), so results do not depend on how many ranks the envs were split over.
"""
import ctypes as C

from . import capi

KIND_RACE, KIND_SWARM = 0, 1
LOG_SUMS = 16  # int64 words of b2d_vec_log_begin


def shard_range(total_envs, rank, world):
    """Contiguous slice [lo, lo+n) of `total_envs` owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


def env_id_base(envs_per_rank, rank):
    """Weak scaling (fixed envs per GPU): global id of this rank's env 0."""
    return int(envs_per_rank) * int(rank)


def reduce_log_sums(sums, group=None):
    """In-place SUM all-reduce of a rank's int64 episode sums (device or host tensor)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def average_log(sums, kind=KIND_RACE, max_rings=10):
    """The C ABI's averaging (b2d_log_average == the tail of b2d_vec_log_end) on host sums:
    9 floats in Log field order, all zeros when no episode finished."""
    vals = [int(x) for x in sums]
    if len(vals) != LOG_SUMS:
        raise ValueError(f"expected {LOG_SUMS} sums")
    arr = (C.c_longlong * LOG_SUMS)(*vals)
    out = (C.c_float * 9)()
    capi.check(capi.lib().b2d_log_average(int(kind), int(max_rings), arr, LOG_SUMS, out))
    return [float(x) for x in out]
