import sys
sys.path[:0]=['.','oracle','tests']
import numpy as np, parity
from continiousenvironment_follower_leader_b200 import capi
d, meta = parity.load_trace('tests/golden/auto_nobear_seed5_follow.npz')
gc = parity.config_for(meta, _route_len=len(d["scen_route"]), _n_static=len(d["scen_static_rects"]))
import os
for libp in [None]+[os.path.abspath('tools/libftl_%s.so'%k) for k in ('ptxO2','ptxO1','regs128','rinit')]:
  for n in (33,):
    env = capi.HostEnv(gc, n, lib=capi.load(libp))
    env.upload_scenarios(parity.pool_for(d, gc))
    env.reset(scenario_ids=np.zeros(n, np.int32))
    a = np.repeat(d["actions"][0][None,:], n, 0)
    out = env.step(a)
    st = env.get_state()
    print(libp, "n=%d" % n, "reward", out.reward[[0, n-1]], "state last_reward", st.env["last_reward"][[0,n-1]], "overall", st.env["overall_reward"][[0,n-1]], "step_count", st.env["step_count"][[0,n-1]], "lpx", st.env["leader"]["pos"][[0,n-1],0])
    env.close()
