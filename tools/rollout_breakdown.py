"""Where the device rollout's time goes: the fused policy kernel alone, the simulator step (fused sensorPrev output) alone,
and the rollout loop; CUDA events, cfg3 workload, 65 536 envs."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
import bench
from continiousenvironment_follower_leader_b200 import capi
from continiousenvironment_follower_leader_b200.rollout import DeviceRollout


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(reps):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n, T = 65536, 16
gc = bench.workload_config(True, fused_sensor_prev=True)
pool, _ = bench.workload_pool(gc)
ro = DeviceRollout(n, T, game_config=gc, scenario_pool=pool)
for _ in range(12):
    ro.collect()
env, st = ro.env, ro.env._stream()
pol = timed(lambda k: ro._fused_policy(ro.obs[k % T], ro._noise_ring[k % T], ro.actions[k % T], ro.values[k % T], st), 200)
sim = timed(lambda k: capi.check(env._L, env._L.ftl_step(env._h, ro.actions[k % T].data_ptr(), C.byref(ro._outs[k % T]), st), "s"), 200)
loop = timed(lambda k: ro.collect(), 10) / T
print("policy kernel %.4f ms   simulator step (fused sensorPrev rows) %.4f ms   rollout loop %.4f ms per step -> %.1f M env-steps/s"
      % (pol, sim, loop, n / loop / 1e3))
ro.close()
