"""tests/golden/multi_end_points.json from the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY (needs /root/reference).

Game(multiple_end_points=True).reset() draws three finish points and chains three D* runs (ENV:471-482, 1552-1611).  The
fixture keeps, per seed, what does not depend on D*'s address-dependent tie-breaks: the finish points, the leader's start,
the number of waypoints, the first waypoint and the flag.

    python oracle/gen_multi_end_golden.py
"""
import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as R

SEEDS = [1, 2, 3, 4, 6, 8]
out = {"env_id": "Test-Cont-Env-Auto-v0", "kwargs": {"multiple_end_points": True}, "cases": []}
for seed in SEEDS:
    env = R.make_env(out["env_id"], **out["kwargs"])
    random.seed(seed)
    np.random.seed(seed)
    with R.quiet():
        env.reset()
    out["cases"].append({
        "seed": seed,
        "finish_points": [list(map(int, p)) for p in (env.finish_point, env.finish_point2, env.finish_point3)],
        "leader_start": [int(v) for v in env.leader.start_position],
        "n_route": len(env.trajectory),
        "route_first": [int(v) for v in env.trajectory[0]],
        "route": [[int(v) for v in p] for p in env.trajectory],
        "found_target_point": bool(env.found_target_point),
    })
    print(seed, out["cases"][-1]["finish_points"], out["cases"][-1]["n_route"], out["cases"][-1]["found_target_point"])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "multi_end_points.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", os.path.normpath(path))
