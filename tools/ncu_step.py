"""Driver for ncu captures: the bench workload (cfg3, 65 536 envs, reference pool), `settle` untimed steps, then `steps`
more -- so that `--kernel-name regex:^k_ --launch-skip <4 * settle + 8> --launch-count 4` captures one steady-state
launch of each kernel of a step."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv

settle = int(sys.argv[1]) if len(sys.argv) > 1 else 200
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
gc = bench.workload_config(True)
pool, _ = bench.workload_pool(gc)
env = FtlBatchEnv(65536, game_config=gc, scenario_pool=pool)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
acts = (lo + (hi - lo) * torch.rand((16, 65536, 2), generator=g, device="cuda")).contiguous()
for k in range(settle + steps):
    env.step_raw(acts[k % 16])
torch.cuda.synchronize()
print("launches", env.launch_count)
env.close()
