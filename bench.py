#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched "follow the leader" simulator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs-per-gpu E]

One "step" = one Game.step for every env of the batch: frames_per_step physics frames + one sensor
scan (reference: follow_the_leader_continuous_env.py:908-945).  Workload at every N: BASELINE.json
configs[2] per GPU -- 65536 envs, 35 static + 1 dynamic obstacles, LeaderPositionsTracker_v2 +
LeaderCorridor_Prev_lasers_v2 (12 rays) + LaserPrevSensor (36 rays), H = 5, F = 10 -- i.e. weak scaling,
each rank owns a disjoint env slice keyed by global env id; NCCL is used only to all-reduce the episode
statistics vector every 64 steps.

The printed JSON line carries
  value     device-timed whole-job env-steps/s with actions resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call (ftl_step_host): actions copied from pinned
            host memory and observation/reward/done copied back to pinned host memory every step
  roofline  algorithmic bytes of the dominant kernel / its CUDA-event time, against the measured HBM peak
  cpu_baseline  the CPU oracle (a C port of the reference's algorithm, oracle/ftl_oracle.c) on the box's
            host cores, bounded sample
`--impl reference` times that CPU oracle alone (the reference itself is pure Python + pygame and cannot
travel to the GPU box; its measured speed in the build container is quoted in BASELINE.md/DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors  # noqa: E402
from continiousenvironment_follower_leader_b200.scenario import ScenarioPool, synthetic_pool  # noqa: E402

METRIC = "env-steps/sec (device-timed, whole box)"
UNIT = "env-steps/s"
WORKLOAD = ("cfg3: 65536 envs/GPU, 35 static + 1 dynamic obstacles, LeaderPositionsTracker_v2 + "
            "LeaderCorridor_Prev_lasers_v2(12 rays) + LaserPrevSensor(36 rays), H=5, frames_per_step=10")

# Algorithmic bytes per env-step (DESIGN.md "Measurement"; SURVEY.md section 8(d) gives 4.3 KB for both kernels)
BYTES_STEP_KERNEL = 2400   # robot/bookkeeping state R+W, static rect table, green window, action, small outputs
BYTES_RAY_KERNEL = 2500    # corridor ring, static rect table, dynamic snapshots, 240 ray floats out
FLOP_PER_ENV_STEP = 3.9e5  # SURVEY.md section 8(d), de-duplicated ray casting + F * controllers


def workload_config(auto_reset=True):
    return GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=auto_reset)


def workload_pool(gc, n_scenarios=1024, seed=0):
    ref = os.path.join(ROOT, "continiousenvironment_follower_leader_b200", "data", "pool_cfg3_reference.npz")
    if os.path.exists(ref):   # scenarios exported from the reference's own reset() (oracle/gen_scenario_pool.py)
        d = np.load(ref)
        if d["static_rects"].shape[1] == gc.c.static_cap and d["route"].shape[1] == gc.c.route_cap:
            return ScenarioPool.from_arrays(d), "reference-reset"
    return synthetic_pool(gc, n_scenarios, seed=seed), "synthetic"


def measured_traffic():
    """DRAM bytes per launch of each kernel from the committed ncu --set full capture (profiles/r01_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------
def cpu_oracle_throughput(gc, pool, budget_s=12.0, n_envs=None, min_steps=4, warmup=1):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle_py import OracleEnv
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    n = n_envs or max(256, 128 * cores)
    env = OracleEnv(gc, n, n_threads=cores)
    env.upload_scenarios(pool)
    env.reset()
    rng = np.random.RandomState(0)
    lo, hi = gc.action_bounds()
    acts = [rng.uniform(lo, hi, size=(n, 2)).astype(np.float32) for _ in range(8)]
    for w in range(warmup):
        env.step(acts[w % 8])
    t0 = time.perf_counter()
    steps = 0
    per_step = []
    while True:
        t1 = time.perf_counter()
        env.step(acts[steps % 8])
        per_step.append(time.perf_counter() - t1)
        steps += 1
        el = time.perf_counter() - t0
        if steps >= min_steps and el >= budget_s:
            break
    el = time.perf_counter() - t0
    env.close()
    return {"value": n * steps / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d envs x %d steps of the same workload (%.1f s) through oracle/ftl_oracle.c, %d OpenMP threads"
                      % (n, steps, el, cores)}, per_step, n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    gc = workload_config()
    pool, pool_kind = workload_pool(gc)
    cores = os.cpu_count() or 1
    n = max(256, 128 * cores)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle_py import OracleEnv
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    env = OracleEnv(gc, n, n_threads=cores)
    env.upload_scenarios(pool)
    env.reset()
    rng = np.random.RandomState(0)
    lo, hi = gc.action_bounds()
    acts = [rng.uniform(lo, hi, size=(n, 2)).astype(np.float32) for _ in range(8)]
    # bound the run: at most ~150 s of CPU work in total
    t0 = time.perf_counter()
    env.step(acts[0])
    one = time.perf_counter() - t0
    steps = max(1, min(args.steps, int(120.0 / max(one, 1e-6))))
    warm = max(0, min(args.warmup, int(20.0 / max(one, 1e-6))))
    for w in range(warm):
        env.step(acts[w % 8])
    t0 = time.perf_counter()
    for k in range(steps):
        env.step(acts[k % 8])
    el = time.perf_counter() - t0
    value = n * steps / el
    sample = ("%d envs x %d steps per run of the cfg3 workload through oracle/ftl_oracle.c (C port of the reference "
              "algorithm), %d OpenMP threads; the pure-Python reference itself measured 16 env-steps/s/process in the "
              "build container" % (n, steps, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic (seeded scenario pool: %s)" % pool_kind,
            "config": {"workload": WORKLOAD, "envs_in_sample": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU leg
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from continiousenvironment_follower_leader_b200 import capi
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs_per_gpu
    gc = workload_config(auto_reset=True)
    pool, pool_kind = workload_pool(gc)
    env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool, device=dev, env_id_base=rank * n)
    env.reset()

    # actions: uniform in the action Box, generated on the device from a per-rank seed; a ring of 16 batches
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    lo, hi = [torch.tensor(x, device=dev) for x in gc.action_bounds()]
    n_act = 16
    actions = (lo + (hi - lo) * torch.rand((n_act, n, 2), generator=g, device=dev)).contiguous()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stats_every = 64
    for w in range(args.warmup):
        env.step_raw(actions[w % n_act])
    if world > 1:
        env.stats(reduce_across_ranks=True)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = env.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        env.step_raw(actions[k % n_act])
        if world > 1 and (k + 1) % stats_every == 0:
            env.stats(reduce_across_ranks=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count - launches0
    # per-kernel split from a separate pass with CUDA events around each kernel on the launching stream: the events
    # switch off the overlap of the ray kernel with the step kernel's tail, so these add up to more than ms_per_step
    env.profile(True)
    for k in range(min(args.steps, 100)):
        env.step_raw(actions[k % n_act])
    step_ms, rays_ms, prof_steps = env.profile_read()
    env.profile(False)
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI (pinned host memory both ways) ---------------------
    # (1) one synchronous ftl_step_host per step over the whole batch; (2) the batch as two halves driven alternately
    # with ftl_step_host_begin / _wait on handle-owned streams (what a double-buffered rollout loop does: while one
    # half's observations cross PCIe -- and its policy would run -- the other half's kernels execute).  Every step of
    # every env still pays its H2D action copy and the D2H copy of all outputs.  (2) is the headline e2e.
    host_actions = [actions[k].cpu().numpy() for k in range(4)]

    def all_max(seconds):
        tt = torch.tensor([seconds], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_sync_value = None
    if args.e2e_steps > 0:
        host = capi.HostEnv(gc, n, device=local_rank, env_id_base=rank * n, pinned=True)
        host.upload_scenarios(pool)
        host.reset()
        for w in range(3):
            host.step(host_actions[w % 4])
        barrier()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            host.step(host_actions[k % 4])
        torch.cuda.synchronize(dev)
        e2e_sync_value = world * n * args.e2e_steps / all_max(time.perf_counter() - t0)
        host.close()

    e2e_value, h2d, d2h = None, 0, 0
    if args.e2e_steps > 0:
        half = n // 2
        parts = [capi.HostEnv(gc, m, device=local_rank, env_id_base=rank * n + first, pinned=True, own_stream=True)
                 for first, m in ((0, half), (half, n - half))]
        acts = [[a[:half] for a in host_actions], [a[half:] for a in host_actions]]
        for hpart in parts:
            hpart.upload_scenarios(pool)
            hpart.reset()
        A, B = parts

        def run(steps):
            A.step_begin(acts[0][0])
            for k in range(steps):
                B.step_begin(acts[1][k % 4])
                A.step_wait()                      # A's observation of step k is in host memory here
                if k + 1 < steps:
                    A.step_begin(acts[0][(k + 1) % 4])
                B.step_wait()

        run(3)
        barrier()
        t0 = time.perf_counter()
        run(args.e2e_steps)
        torch.cuda.synchronize(dev)
        e2e_value = world * n * args.e2e_steps / all_max(time.perf_counter() - t0)
        h2d = sum(x.h2d_bytes_per_step for x in parts)
        d2h = sum(x.d2h_bytes_per_step for x in parts)
        for hpart in parts:
            hpart.close()

    stats = env.stats_dict(reduce_across_ranks=world > 1)
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # k_rays includes the (almost always idle) exact-resolution launch k_rays_exact that follows it
        k_ms = {"k_step": step_ms / max(prof_steps, 1), "k_rays": rays_ms / max(prof_steps, 1)}
        dominant = max(k_ms, key=k_ms.get)
        alg_bytes = (BYTES_STEP_KERNEL if dominant == "k_step" else BYTES_RAY_KERNEL) * n
        achieved = alg_bytes / (k_ms[dominant] * 1e-3) / 1e9
        traffic = measured_traffic().get(dominant)
        if traffic is not None and n != 65536:
            traffic = None   # the capture was taken at 65536 envs per GPU
        clocks = sampler.summary()
        sm_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 controllers + f32 positions/rays + i32 hitboxes", "data": "synthetic (seeded scenario pool: %s, %d scenarios; "
            "uniform random actions)" % (pool_kind, pool.n),
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "global_envs": n * world, "auto_reset": True,
                       "cache": "state touched per step (%.0f MB) exceeds the 126 MB L2; no explicit flush" % (
                           n * (BYTES_STEP_KERNEL + BYTES_RAY_KERNEL) / 1e6),
                       "stats_allreduce_every": stats_every if world > 1 else None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps,
                    "api": "two capi.HostEnv halves, step_begin/step_wait alternating -> ftl_step_host_begin/_wait "
                           "(pinned host buffers, handle-owned streams)",
                    "synchronous_whole_batch": e2e_sync_value,
                    "synchronous_api": "capi.HostEnv.step -> ftl_step_host, one call per step for all envs"},
            "gpu_launches": int(launches),
            "kernels_ms_per_step": k_ms,
            "kernels_note": "each kernel timed alone (events between the launches disable the overlap): k_rays is a "
                            "programmatic dependent launch of k_step and fills its tail, so ms_per_step < k_step + k_rays",
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_kind,
                         "note": "k_rays is issue-bound (62% of issue slots busy, ncu), not HBM-bound: measured DRAM traffic is 1.3 KB/env-step",
                         "algorithmic_bytes_per_env_step": BYTES_STEP_KERNEL if dominant == "k_step" else BYTES_RAY_KERNEL,
                         "fp32_alu": {"flop_per_env_step": FLOP_PER_ENV_STEP,
                                      "achieved_tflops": FLOP_PER_ENV_STEP * n * args.steps / (ms * 1e-3) / 1e12,
                                      "peak_tflops_at_measured_clock": fp32_peak}},
            "episode_stats": stats,
        }
        if not args.no_cpu_baseline and world == 1:
            cb, _, _ = cpu_oracle_throughput(gc, pool)
            line["cpu_baseline"] = cb
        elif world > 1:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
