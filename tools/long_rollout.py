import sys, torch
sys.path.insert(0, ".")
import bench
from continiousenvironment_follower_leader_b200.rollout import DeviceRollout
gc = bench.workload_config(True, fused_sensor_prev=True)
pool, _ = bench.workload_pool(gc)
ro = DeviceRollout(65536, 64, game_config=gc, scenario_pool=pool)
tot = 0
for k in range(40):
    traj = ro.collect(explore=True)
    adv, ret = ro.advantages()
    tot += 64
    if k % 10 == 9:
        torch.cuda.synchronize()
        print("steps", tot, "obs finite", bool(torch.isfinite(traj["obs"]).all()), "act range", float(traj["actions"].min()), float(traj["actions"].max()),
              "mean reward %.4f" % float(traj["rewards"].mean()), "done rate %.5f" % float(traj["dones"].float().mean()), "adv finite", bool(torch.isfinite(adv).all()), flush=True)
st = ro.env.stats_dict()
print({k: st[k] for k in list(st)[:5]})
ro.close()
