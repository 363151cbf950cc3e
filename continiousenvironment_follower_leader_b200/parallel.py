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


def gpu_numa_node(gpu_index):
    """NUMA node of a GPU from sysfs (PCI bus id through NVML), or None when the platform does not say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:       # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _cpus_of_node(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            text = f.read().strip()
    except Exception:
        return None
    cpus = set()
    for part in text.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa(gpu_index, local_world=None):
    """Pin the calling process to the cores of the NUMA node nearest `gpu_index` (so that pinned host buffers
    allocated AFTERWARDS land on that node and the D2H copies of the host path do not cross sockets).  When several
    ranks share a node its cores are split evenly between them.  Returns what was done; never raises."""
    info = {"gpu": int(gpu_index), "numa_node": None, "bound": False}
    try:
        allowed = os.sched_getaffinity(0)
        node = gpu_numa_node(gpu_index)
        info["numa_node"] = node
        cpus = _cpus_of_node(node) if node is not None else None
        if not cpus:
            return info
        cpus = sorted(cpus & allowed)
        if not cpus:
            return info
        if local_world is None:
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
        # ranks whose GPUs sit on the same node share its cores
        same = [g for g in range(local_world) if gpu_numa_node(g) == node] or [int(gpu_index)]
        k = same.index(int(gpu_index)) if int(gpu_index) in same else 0
        share = max(1, len(cpus) // len(same))
        mine = cpus[k * share:(k + 1) * share] or cpus
        os.sched_setaffinity(0, set(mine))
        info.update(bound=True, cpus=mine)
    except Exception as e:   # noqa: BLE001 - binding is an optimisation
        info["error"] = repr(e)
    return info


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
