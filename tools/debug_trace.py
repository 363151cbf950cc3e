"""Replay a golden trace through libftl.so and the oracle side by side; print the first differing leader fields."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import numpy as np
import parity
from continiousenvironment_follower_leader_b200 import capi
from oracle_py import OracleEnv

name = sys.argv[1]
d, meta = parity.load_trace(parity.GOLDEN_DIR + "/%s.npz" % name)
gc = parity.config_for(meta, _route_len=len(d["scen_route"]), _n_static=len(d["scen_static_rects"]))
n = 33
cuda, orc = capi.HostEnv(gc, n, lib=capi.load()), OracleEnv(gc, n)
pool = parity.pool_for(d, gc)
cuda.upload_scenarios(pool); orc.upload_scenarios(pool)
ids = np.zeros(n, np.int32)
cuda.reset(scenario_ids=ids); orc.reset(scenario_ids=ids)
for t, a in enumerate(d["actions"]):
    A = np.repeat(a[None, :], n, axis=0)
    kw = {}
    if "step_frames" in d:
        kw = dict(frames=np.full(n, d["step_frames"][t], np.int32), regime_draws=np.repeat(d["step_draws"][t][None, :], n, axis=0))
    cuda.step(A, **kw); orc.step(A, **kw)
    sc, so = cuda.get_state().env[0], orc.get_state().env[0]
    bad = []
    for f in ("cur_speed_multiplier", "cur_leader_acceleration", "cur_leader_cumulative_speed", "accel_consumed", "step_count"):
        if sc[f] != so[f]:
            bad.append((f, sc[f], so[f]))
    for f in ("speed", "des_speed", "dir"):
        if abs(sc["leader"][f] - so["leader"][f]) > 1e-9:
            bad.append(("leader." + f, sc["leader"][f], so["leader"][f]))
    if bad or t in (45, 46, 47):
        print("step", t, "frames", kw.get("frames", [None])[0], "step_count", so["step_count"], bad[:6])
    if bad:
        break
