"""N > 1 on the CPU: two gloo ranks, each owning half of the env batch (host build of the device functions).

Checks what the multi-GPU path relies on: slices are disjoint and keyed by global env id, so the two halves
reproduce the single-process batch bit for bit, and the all-reduced statistics equal the single-process ones.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(gc_kwargs, base, count, steps, seed_pool=4):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from continiousenvironment_follower_leader_b200 import abi, parallel
    from continiousenvironment_follower_leader_b200.config import GameConfig
    from continiousenvironment_follower_leader_b200.scenario import synthetic_pool
    from hostsim_py import make_env
    gc = GameConfig(**gc_kwargs)
    pool = synthetic_pool(gc, 16, seed=seed_pool)
    env = make_env(gc, count, env_id_base=base)
    env.upload_scenarios(pool)
    env.reset()
    rng = np.random.RandomState(99)
    lo, hi = gc.action_bounds()
    total = 64
    stats = torch.zeros(abi.STAT_COUNT, dtype=torch.float64)
    nf, rays = [], []
    for t in range(steps):
        a_all = rng.uniform(lo, hi, size=(total, 2)).astype(np.float32)      # same global action stream everywhere
        before = env.get_state().env["done"].copy()
        out = env.step(a_all[base:base + count])
        # statistics of finished episodes must be read before the in-place auto-reset; the kernel does this on the
        # device, here they come from the outputs (reward of the last frame is not the return, so use done only)
        stats[abi.STAT_EPISODES] += float(out.done.sum())
        stats[abi.STAT_CRASH] += float((out.status[:, 3] != 0)[out.done.astype(bool)].sum())
        nf.append(out.numerical_features.copy())
        rays.append(out.rays.copy())
    return np.stack(nf), np.stack(rays), stats


def _worker(rank, world, port, gc_kwargs, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from continiousenvironment_follower_leader_b200 import parallel
    base, count = parallel.shard(64, rank, world)
    nf, rays, stats = _run(gc_kwargs, base, count, steps)
    stats = parallel.reduce_stats(stats)
    q.put((rank, base, count, nf, rays, stats.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_batch():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from continiousenvironment_follower_leader_b200.config import cfg3_sensors
    gc_kwargs = dict(bear_number=1, follower_sensors=cfg3_sensors(), max_steps=150, auto_reset=True)
    steps = 40
    ref_nf, ref_rays, ref_stats = _run(gc_kwargs, 0, 64, steps)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gc_kwargs, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ref_stats[0] > 64          # every env finished at least once: auto-reset and re-keyed scenarios exercised
    for rank, base, count, nf, rays, stats in results:
        assert np.array_equal(nf, ref_nf[:, base:base + count]), "rank %d slice differs" % rank
        assert np.array_equal(rays, ref_rays[:, base:base + count])
        assert np.array_equal(stats, ref_stats.numpy()), "all-reduced statistics differ"
    assert sorted(r[1] for r in results) == [0, 32]


def test_shard_covers_the_batch_without_overlap():
    from continiousenvironment_follower_leader_b200 import parallel
    for world in (1, 2, 3, 4, 8):
        spans = [parallel.shard(1000, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == 1000
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1
