"""ms per step of the bench workload (cfg3, 65 536 envs, reference pool) as a function of the step index since the common
reset: every env starts its first episode together, so the mix of episode ages (trail lengths, resets per step) only
becomes stationary after several hundred steps (an episode lasts at most 501 steps).  CUDA events per chunk of steps."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 100
total = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
gc = bench.workload_config(True)
pool, _ = bench.workload_pool(gc)
n = 65536
env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
acts = (lo + (hi - lo) * torch.rand((16, n, 2), generator=g, device="cuda")).contiguous()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(total // chunk + 1)]
evs[0].record()
for c in range(total // chunk):
    for k in range(chunk):
        env.step_raw(acts[(c * chunk + k) % 16])
    evs[c + 1].record()
torch.cuda.synchronize()
st = env.stats_dict()
for c in range(total // chunk):
    ms = evs[c].elapsed_time(evs[c + 1]) / chunk
    print("steps %5d-%5d  %.4f ms per step  %.1f M env-steps/s" % (c * chunk, (c + 1) * chunk, ms, n / ms / 1e3))
print({k: st[k] for k in list(st)[:4]})
env.close()
