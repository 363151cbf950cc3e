"""Device-time of the two kernels versus frames_per_step / ray count / env count (diagnostic sweep)."""
import os, sys, json
import torch
sys.path.insert(0, ".")
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors
from continiousenvironment_follower_leader_b200.scenario import synthetic_pool

def run(n, F, rays=(12, 36), steps=200, warm=100, auto_reset=True, max_steps=5000, ref_pool=False, checksum=False):
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(rays[0], rays[1]), frames_per_step=F, auto_reset=auto_reset, max_steps=max_steps)
    pool = synthetic_pool(gc, 256, seed=0)
    if ref_pool:
        import numpy as np
        from continiousenvironment_follower_leader_b200.scenario import ScenarioPool
        pool = ScenarioPool.from_arrays(np.load('continiousenvironment_follower_leader_b200/data/pool_cfg3_reference.npz'))
    env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool, lib_path=os.environ.get('AB_LIB'))   # AB_LIB: a tools/libftl_<tag>.so variant (tools/ab_libs.py)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
    acts = lo + (hi - lo) * torch.rand((8, n, 2), generator=g, device="cuda")
    for k in range(warm): env.step_raw(acts[k % 8])
    env.profile(True)
    for k in range(steps): env.step_raw(acts[k % 8])
    km, c = env.profile_read_kernels()
    a, b = km["k_kin"] + km["k_book"], km["k_rays"]
    run.last_kernels = {k: v / c for k, v in km.items()}
    env.profile(False)
    # whole steps without per-kernel events: the ray kernel may overlap the tail of the step kernel
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(20): env.step_raw(acts[k % 8])
    e0.record()
    for k in range(steps): env.step_raw(acts[k % 8])
    e1.record()
    torch.cuda.synchronize()
    run.last_total_ms = e0.elapsed_time(e1) / steps
    h = None
    if checksum:
        import hashlib
        st = env.get_state()
        m = hashlib.sha1()
        for f in ("step_count", "trail_len", "cur_target_id", "saving_counter"):
            m.update(st.env[f].tobytes())
        m.update(st.env["follower"]["rect"].tobytes())
        m.update(env.rays.cpu().numpy().tobytes())
        h = m.hexdigest()[:12]
    env.close()
    return (a / c, b / c, h) if checksum else (a / c, b / c)

if __name__ == "__main__":
    # BASELINE.json configs[4]: frames_per_step sweep and ray-count sweep (single GPU; 262144 envs over 8 GPUs = 32768 per GPU)
    print("== frames_per_step sweep, 32768 envs, cfg3 sensors")
    for F in (2, 3, 5, 8, 10):
        a, b = run(32768, F, ref_pool=True)
        print("F=%2d  k_step %.4f ms  k_rays %.4f ms  step %.4f ms -> %.1f M env-steps/s" % (F, a, b, run.last_total_ms, 32768 / run.last_total_ms / 1e3), flush=True)
    print("== obstacle-sensor ray-count sweep, 32768 envs, F=10")
    for R in (12, 24, 36, 72, 120, 180, 360):
        a, b = run(32768, 10, rays=(12, R), ref_pool=True)
        print("R=%3d  k_step %.4f ms  k_rays %.4f ms  step %.4f ms -> %.1f M env-steps/s" % (R, a, b, run.last_total_ms, 32768 / run.last_total_ms / 1e3), flush=True)
    print("== env-count sweep, F=10, cfg3")
    for n in (4096, 16384, 65536, 131072, 262144):
        a, b = run(n, 10, ref_pool=True)
        print("N=%6d  k_step %.4f ms  k_rays %.4f ms  step %.4f ms -> %.1f M env-steps/s" % (n, a, b, run.last_total_ms, n / run.last_total_ms / 1e3), flush=True)
