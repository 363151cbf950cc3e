"""Export a pool of scenarios drawn by the UNMODIFIED reference's own reset() -- build container only.

    python oracle/gen_scenario_pool.py [n_scenarios]

Writes continiousenvironment_follower_leader_b200/data/pool_cfg3_reference.npz: layouts (walls + 35 rocks),
D* routes, leader/follower start poses for BASELINE.json configs[2], seeds 1, 2, 3, ... with unreachable-route
seeds skipped the way SkipBadSeeds does (utils/wrappers.py:814-825).  bench.py and the tests use this pool
so that the benchmark workload has the reference's scenario distribution.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.dirname(_HERE))


def _one(seed):
    import ref_harness as rh
    import gen_golden
    env = rh.make_env("Test-Cont-Env-Auto-v0", bear_number=1)
    rh.reset(env, seed=seed)
    if not env.found_target_point or len(env.trajectory) > 126:
        return seed, None
    return seed, gen_golden.extract_scenario(env)


def main(n_want=512):
    from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors
    from continiousenvironment_follower_leader_b200.scenario import ScenarioPool
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors())
    got = []
    seed = 1
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        while len(got) < n_want:
            batch = list(range(seed, seed + 64))
            seed += 64
            for s, sc in pool.map(_one, batch):
                if sc is not None and len(got) < n_want:
                    got.append((s, sc))
            print("seeds < %d: %d scenarios" % (seed, len(got)), flush=True)
    sp = ScenarioPool(len(got), gc.c.static_cap, gc.c.route_cap)
    for i, (s, sc) in enumerate(got):
        sp.set(i, sc["static_rects"], sc["route"], sc["leader_pos"], float(sc["leader_dir"]), sc["follower_pos"],
               float(sc["follower_dir"]), True)
    out = os.path.join(os.path.dirname(_HERE), "continiousenvironment_follower_leader_b200", "data",
                       "pool_cfg3_reference.npz")
    sp.save(out)
    np.save(out.replace(".npz", "_seeds.npy"), np.array([s for s, _ in got], np.int32))
    print("wrote", out, os.path.getsize(out) // 1024, "KB")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 512)
