"""Diagnostic for A/B log (24): a GPU build with a 24-edge list against the oracle; the first env-step whose rays differ
is replayed (state before the step + its action) through instrumented builds that print the ray pass's edge counters."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import numpy as np
import parity
from continiousenvironment_follower_leader_b200 import capi
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors
from continiousenvironment_follower_leader_b200.scenario import synthetic_pool
from oracle_py import OracleEnv

kwargs = dict(bear_number=1, frames_per_step=60, follower_sensors=cfg3_sensors(), max_steps=2000, auto_reset=True)
n, steps = 512, 40
gc = GameConfig(**kwargs)
pool = synthetic_pool(gc, 64, seed=2)
cuda, orc = capi.HostEnv(gc, n, lib=capi.load("tools/libftl_sc_a.so")), OracleEnv(gc, n, n_threads=8)
cuda.upload_scenarios(pool); orc.upload_scenarios(pool)
ids = (np.arange(n) % pool.n).astype(np.int32)
cuda.reset(scenario_ids=ids); orc.reset(scenario_ids=ids)
rng = np.random.RandomState(17)
found = None
for t in range(steps):
    a = parity.sample_actions(gc, rng, n, t)
    before = cuda.get_state()
    oc, oo = cuda.step(a), orc.step(a)
    w = np.argwhere(~np.isclose(oc.rays, oo.rays, rtol=1e-4, atol=1e-4))
    if len(w):
        e = int(w[0, 0])
        print("step", t, "env", e, "outliers", int((w[:, 0] == e).sum()), flush=True)
        found = (before, e, a[e].copy(), oo.rays[e].copy(), oc.rays[e].copy())
        break
assert found, "no outlier found"
before, e, act, want, got = found
m = 32
for tag in ("dbg_small", "dbg_default"):
    env = capi.HostEnv(gc, m, lib=capi.load("tools/libftl_%s.so" % tag))
    env.upload_scenarios(pool)
    env.reset(scenario_ids=np.full(m, ids[e], np.int32))
    st = env.get_state()
    for name in ("env", "trail", "hist", "corridor"):
        getattr(st, name)[:] = getattr(before, name)[e]
    env.set_state(st)
    for rep in range(3):
        env.set_state(st)
        o = env.step(np.tile(act, (m, 1)))
        bad = [int((~np.isclose(o.rays[k], want, rtol=1e-4, atol=1e-4)).sum()) for k in range(m)]
        print(tag, "rep", rep, "bad rays per copy:", bad[:8], "max", max(bad), flush=True)
    env.close()
