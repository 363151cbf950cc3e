"""Concurrent pinned D2H ceiling of the box: N ranks each copy `--mb` MB device -> pinned host, no kernels.

    python tools/d2h_ceiling.py --gpus 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/d2h_ceiling.py --gpus 8

Prints one JSON line (rank 0): per-rank and aggregate GB/s, first with the ranks floating over all cores, then with
each rank bound to the cores / NUMA node nearest its GPU (parallel.bind_to_gpu_numa; the pinned buffer is allocated
after binding).  bench.py's e2e leg moves 66.65 MB of observations per step per GPU over this path; the ratio
e2e bytes/s / ceiling says whether the e2e number is the host's limit or the code's.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from continiousenvironment_follower_leader_b200 import parallel  # noqa: E402


def measure(dev, mb, reps, world):
    n = int(mb * 1e6) // 4
    src = torch.empty(n, dtype=torch.float32, device=dev).normal_()
    dst = torch.empty(n, dtype=torch.float32).pin_memory()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    el = time.perf_counter() - t0
    return n * 4 * reps / el / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mb", type=float, default=66.65)
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {}
    for label, bind in (("unbound", False), ("bound", True)):
        info = parallel.bind_to_gpu_numa(local) if bind else {"cpus": len(os.sched_getaffinity(0))}
        gbs = measure(dev, args.mb, args.reps, world)
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, t)
        else:
            allv = [t]
        vals = [float(x.item()) for x in allv]
        if isinstance(info.get("cpus"), list):
            info = dict(info, cpus="%d cores: %s.." % (len(info["cpus"]), info["cpus"][:4]))
        out[label] = {"per_rank_gbs": vals, "aggregate_gbs": sum(vals), "min_gbs": min(vals), "rank0_binding": info}
    if rank == 0:
        print(json.dumps({"tool": "d2h_ceiling", "n_gpus": world, "mb_per_copy": args.mb, "reps": args.reps, **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
