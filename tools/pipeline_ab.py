"""Device time per 65 536 env-steps with the batch driven as 1, 2 or 4 independent handles on their own CUDA streams
(every handle's step t+1 is enqueued behind its own step t; the handles alternate, so one handle's ray kernel overlaps
another handle's latency-bound kinematics / bookkeeping kernels).  Diagnostic for bench.py's pipelined leg."""
import os, sys
import torch
sys.path.insert(0, ".")
import bench
from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv


def run(parts, n_total=65536, steps=200, settle=150):
    gc = bench.workload_config(True)
    pool, _ = bench.workload_pool(gc)
    n = n_total // parts
    envs = [FtlBatchEnv(n, game_config=gc, scenario_pool=pool, env_id_base=k * n, lib_path=os.environ.get("AB_LIB")) for k in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    g = torch.Generator(device="cuda").manual_seed(1234)
    lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
    acts = (lo + (hi - lo) * torch.rand((16, n_total, 2), generator=g, device="cuda")).contiguous()
    views = [[acts[j, k * n:(k + 1) * n] for j in range(16)] for k in range(parts)]
    torch.cuda.synchronize()
    for e, st in zip(envs, streams):
        with torch.cuda.stream(st):
            e.reset()

    def go(count):
        for t in range(count):
            for k, (e, st) in enumerate(zip(envs, streams)):
                with torch.cuda.stream(st):
                    e.step_raw(views[k][t % 16])

    go(settle)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(parts)]
    e0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(e0)
    go(steps)
    for ev, st in zip(ends, streams):
        ev.record(st)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in ends) / steps
    for e in envs:
        e.close()
    return ms


if __name__ == "__main__":
    for parts in (1, 2, 4):
        ms = run(parts)
        print("handles %d x %6d envs: %.4f ms per 65536 env-steps -> %.1f M env-steps/s" % (parts, 65536 // parts, ms, 65536 / ms / 1e3), flush=True)
