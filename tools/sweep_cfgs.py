"""Device time per step of the other BASELINE.json configurations (diagnostic; parity for them is in tests/):
cfg2 (4096 envs, no obstacles, tracker only) and cfg4 (the "gazebo" preset, hardcore physics, at 131072 envs per GPU)."""
import sys
import torch
sys.path.insert(0, ".")
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors, TEST_GAME_MANUAL_GAZEBO_KWARGS
from continiousenvironment_follower_leader_b200.scenario import synthetic_pool


def run(name, kwargs, n, steps=200, warm=100):
    gc = GameConfig(**kwargs)
    env = FtlBatchEnv(n, game_config=gc, scenario_pool=synthetic_pool(gc, 256, seed=0))
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
    acts = lo + (hi - lo) * torch.rand((8, n, 2), generator=g, device="cuda")
    for k in range(warm): env.step_raw(acts[k % 8])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps): env.step_raw(acts[k % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    env.close()
    print("%-28s N=%7d  F=%2d  step %.4f ms -> %.1f M env-steps/s" % (name, n, gc.c.frames_per_step, ms, n / ms / 1e3), flush=True)


if __name__ == "__main__":
    tracker = {"LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"]}
    run("cfg2 (no obstacles)", dict(add_obstacles=False, add_bear=False, follower_sensors=tracker, auto_reset=True), 4096)
    run("cfg2 at 65536 envs", dict(add_obstacles=False, add_bear=False, follower_sensors=tracker, auto_reset=True), 65536)
    run("cfg4 (gazebo preset)", dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, auto_reset=True), 131072)
    run("cfg4 at 65536 envs", dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, auto_reset=True), 65536)
