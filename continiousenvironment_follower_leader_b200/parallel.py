"""Multi-GPU plumbing: the env batch shards trivially, one process per GPU.

Each rank owns the disjoint env slice [base, base + count) of the global batch; everything that is random
or scenario-dependent inside the kernels is keyed by the GLOBAL env id (``env_id_base`` of ftl_create), so
results do not depend on how many ranks the batch is split over.  The only collective is a sum of the
episode-statistics vector (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def shard(global_envs, rank=None, world=None):
    """(base, count) of this rank's slice; slices differ by at most one env."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    q, r = divmod(int(global_envs), world)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def reduce_stats(stats):
    """Sum a statistics tensor over all ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def episode_stats_from_outputs(done_before, done_after, overall_reward, step_count, status):
    """The statistics vector the step kernel accumulates (include/ftl.h, FTL_STAT_*), from host arrays: used where
    the kernel's own atomics are not available (the CPU test harness)."""
    import numpy as np
    from . import abi
    v = torch.zeros(abi.STAT_COUNT, dtype=torch.float64)
    new = (~done_before.astype(bool)) & done_after.astype(bool)
    v[abi.STAT_EPISODES] = float(new.sum())
    v[abi.STAT_RETURN_SUM] = float(np.asarray(overall_reward)[new].sum())
    v[abi.STAT_LENGTH_SUM] = float(np.asarray(step_count)[new].sum())
    st = np.asarray(status)[new]
    v[abi.STAT_CRASH] = float((st[:, 3] != 0).sum())
    v[abi.STAT_SUCCESS] = float((st[:, 0] == 2).sum())
    v[abi.STAT_TIMEOUT] = float((st[:, 0] == 3).sum())
    v[abi.STAT_LEADER_CRASH] = float((st[:, 2] == 2).sum())
    return v
