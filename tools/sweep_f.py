"""Device-time of the two kernels versus frames_per_step / ray count / env count (diagnostic sweep)."""
import sys, json
import torch
sys.path.insert(0, ".")
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors
from continiousenvironment_follower_leader_b200.scenario import synthetic_pool

def run(n, F, rays=(12, 36), steps=200, warm=100, auto_reset=True, max_steps=5000, ref_pool=False):
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(rays[0], rays[1]), frames_per_step=F, auto_reset=auto_reset, max_steps=max_steps)
    pool = synthetic_pool(gc, 256, seed=0)
    if ref_pool:
        import numpy as np
        from continiousenvironment_follower_leader_b200.scenario import ScenarioPool
        pool = ScenarioPool.from_arrays(np.load('continiousenvironment_follower_leader_b200/data/pool_cfg3_reference.npz'))
    env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
    acts = lo + (hi - lo) * torch.rand((8, n, 2), generator=g, device="cuda")
    for k in range(warm): env.step_raw(acts[k % 8])
    env.profile(True)
    for k in range(steps): env.step_raw(acts[k % 8])
    a, b, c = env.profile_read()
    env.close()
    return a / c, b / c

if __name__ == "__main__":
    for F in (2, 5, 10, 20):
        a, b = run(65536, F)
        print("N=65536 F=%2d  k_step %.4f ms  k_rays %.4f ms" % (F, a, b), flush=True)
    for n in (16384, 32768, 131072, 262144):
        a, b = run(n, 10)
        print("N=%6d F=10  k_step %.4f ms  k_rays %.4f ms  -> %.1f M env-steps/s" % (n, a, b, n / (a + b) / 1e3), flush=True)
    a, b = run(65536, 10, auto_reset=False)
    print("no auto reset: k_step %.4f k_rays %.4f" % (a, b))
