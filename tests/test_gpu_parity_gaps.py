"""GPU parity, part 2 (`pytest -m gpu`): the sub-paths of Game.step that round 1 only covered on the CPU build.

  * constant_follower_speed / aggregate_reward (ENV:918-925, 1136), frames_per_step in {1, 2, 13} (the chunk boundary of
    the frame loop and the low end of BASELINE.json's F sweep), 120 rays -- libftl.so against the oracle on seeded batches
  * the episode-statistics vector the step kernel accumulates with atomics (what NCCL reduces) against the same sums
    recomputed from the oracle's outputs, with and without auto-reset
  * the bench workload itself: 65 536 envs on the 512-layout reference pool with auto-reset; a 1 024-env slice is
    re-simulated by the oracle from the same actions and compared after 100 / 200 / 300 steps

Bars as everywhere: integers bit-exact, floats within 1e-4 relative, ray outliers counted.
"""
import os

import numpy as np
import pytest

import parity
from continiousenvironment_follower_leader_b200 import abi, capi
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors, TEST_GAME_MANUAL_GAZEBO_KWARGS
from continiousenvironment_follower_leader_b200.scenario import ScenarioPool, synthetic_pool
from test_gpu_parity import _compare_states, _ray_outliers

pytestmark = pytest.mark.gpu


sample_actions = parity.sample_actions


GAP_CASES = [
    ("const_speed", dict(bear_number=1, constant_follower_speed=True, follower_sensors=cfg3_sensors()), 1024, 60),
    ("aggregate_reward", dict(bear_number=1, aggregate_reward=True, follower_sensors=cfg3_sensors(), max_steps=400,
                              auto_reset=True), 1024, 70),
    ("f1", dict(bear_number=1, frames_per_step=1, follower_sensors=cfg3_sensors()), 1024, 150),
    ("f2_120rays", dict(bear_number=1, frames_per_step=2, follower_sensors=cfg3_sensors(12, 120, 5)), 512, 90),
    ("f13_2bears", dict(bear_number=2, frames_per_step=13, follower_sensors=cfg3_sensors(), max_steps=400,
                        auto_reset=True), 1024, 50),
    # 60 frames per step: the per-frame records no longer fit k_kin's shared memory and go through HBM
    ("f60_records_in_hbm", dict(bear_number=1, frames_per_step=60, follower_sensors=cfg3_sensors(), max_steps=2000,
                                auto_reset=True), 512, 40),
    ("f10_warm0_es", dict(bear_number=1, warm_start=0, early_stopping={"max_distance_coef": 1.3, "low_reward": -60},
                          follower_sensors=cfg3_sensors(), auto_reset=True), 1024, 80),
    # LeaderCorridor_lasers_compas (SEN:1138-1240; cast by the per-env exact pass of k_finish)
    ("compas", dict(parity.load_trace(parity.GOLDEN_DIR + "/compas_seed25.npz")[1]["kwargs"], auto_reset=True,
                    max_steps=300), 1024, 60),
    # LaserSensor (SEN:18-136; k_optional_sensors)
    ("laser_points", dict(parity.load_trace(parity.GOLDEN_DIR + "/laser_sensor_seed27.npz")[1]["kwargs"], auto_reset=True,
                          max_steps=300), 1024, 60),
    ("laser_distances", dict(parity.load_trace(parity.GOLDEN_DIR + "/laser_distances_seed29.npz")[1]["kwargs"],
                             auto_reset=True, max_steps=300), 1024, 60),
]


@pytest.mark.parametrize("name,kwargs,n,steps", GAP_CASES, ids=[c[0] for c in GAP_CASES])
def test_cuda_matches_oracle_on_the_uncovered_step_options(name, kwargs, n, steps):
    _gap_case(kwargs, n, steps, capi.load())


SMALL_LIST_CASES = [c for c in GAP_CASES if c[0] in ("f60_records_in_hbm", "f13_2bears")] + [
    ("gazebo", dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, max_steps=900, auto_reset=True), 1024, 100)]


@pytest.mark.parametrize("name,kwargs,n,steps", SMALL_LIST_CASES, ids=[c[0] for c in SMALL_LIST_CASES])
def test_ray_kernel_overflow_paths_on_the_gpu(name, kwargs, n, steps):
    """csrc/libftl_smalllists.so (build.build_small_lists: 24 edges, 40 pairs, 2 exact-pass records per env) against the
    oracle: every overflow path of the ray pass with really concurrent lanes.  Long corridors (60 frames per step, the
    Gazebo presets) are the cases in which the lanes of a warp once disagreed on "flush the list first?" -- a decision
    read from a shared counter that faster lanes were already incrementing (FTL_UNIFORM_INT in ftl_rays.cuh)."""
    from continiousenvironment_follower_leader_b200 import build
    path = build.SMALL_LISTS_OUT if os.path.exists(build.SMALL_LISTS_OUT) else build.build_small_lists()   # built by build()
    lib = capi.load(path)
    info = lib.ftl_build_info().decode()
    assert "edge_cap=24 " in info and "pair_cap=40 " in info and "scan_wide=1 " in info, info
    assert "edge_cap=176 " in capi.load().ftl_build_info().decode()
    _gap_case(kwargs, n, steps, lib)


def _gap_case(kwargs, n, steps, lib):
    from oracle_py import OracleEnv
    gc = GameConfig(**kwargs)
    pool = synthetic_pool(gc, 64, seed=2)
    cuda, orc = capi.HostEnv(gc, n, lib=lib), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    cuda.reset(scenario_ids=ids)
    orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(17)
    bad, total, dones, laser_bad, laser_total = 0, 0, 0, 0, 0
    for t in range(steps):
        a = sample_actions(gc, rng, n, t)
        oc, oo = cuda.step(a), orc.step(a)
        assert np.array_equal(oc.done, oo.done), "done differs at step %d" % t
        assert np.array_equal(oc.status, oo.status), "status differs at step %d" % t
        assert np.allclose(oc.reward, oo.reward, rtol=1e-5, atol=1e-5), "reward differs at step %d" % t
        assert np.array_equal(oc.leader_target, oo.leader_target)
        assert np.allclose(oc.numerical_features, oo.numerical_features, rtol=parity.RTOL, atol=1e-4)
        bad += _ray_outliers(oc.rays, oo.rays)
        total += oc.rays.size
        dones += int(oo.done.sum())
        if oc.laser is not None:   # discrete sample points: equal up to float rounding, or a sample on a hit-box edge
            laser_bad += _ray_outliers(oc.laser, oo.laser)
            laser_total += oc.laser.size
        if t % 10 == 9 or t == steps - 1:
            _compare_states(cuda.get_state(), orc.get_state(), gc, n, parity.RTOL)
    assert bad <= 2, "%d of %d ray values outside tolerance" % (bad, total)
    assert laser_bad <= 1e-4 * max(laser_total, 1), "%d of %d LaserSensor values outside tolerance" % (laser_bad, laser_total)
    if gc.c.aggregate_reward:   # the returned reward is the running total (ENV:1136), not the last frame's
        live = ~oo.done.astype(bool)
        assert live.sum() > n // 2
        assert np.allclose(oc.reward[live], orc.get_state().env["overall_reward"][live], rtol=1e-5, atol=1e-5)
        assert np.abs(oc.reward).max() > 5, "aggregate rewards should have grown beyond a single frame's reward"
    cuda.close()


def _stats_from_outputs(done_prev, out, state):
    """FTL_STAT_* sums of the envs that finished in this step, from host arrays."""
    new = (~done_prev) & out.done.astype(bool)
    v = np.zeros(abi.STAT_COUNT, np.float64)
    v[abi.STAT_EPISODES] = new.sum()
    v[abi.STAT_RETURN_SUM] = state.env["overall_reward"][new].sum()
    v[abi.STAT_LENGTH_SUM] = state.env["step_count"][new].sum()
    st = out.status[new]
    v[abi.STAT_CRASH] = (st[:, 3] != 0).sum()
    v[abi.STAT_SUCCESS] = (st[:, 0] == 2).sum()
    v[abi.STAT_TIMEOUT] = (st[:, 0] == 3).sum()
    v[abi.STAT_LEADER_CRASH] = (st[:, 2] == 2).sum()
    return v


@pytest.mark.parametrize("auto_reset", [False, True], ids=["no_reset", "auto_reset"])
def test_device_episode_statistics_equal_the_sums_over_the_oracle(auto_reset):
    """ftl_stats (atomics in the step kernel; the vector NCCL all-reduces) against sums over the oracle.  Without
    auto-reset the oracle's state still holds the finished episode's return and length when it is read; with
    auto-reset the env is already re-initialised, so the run uses aggregate_reward (the returned reward of the last
    step IS the episode return, ENV:1136) and counts the steps of each episode (step_count advances by F per step)."""
    import torch
    from oracle_py import OracleEnv
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), max_steps=250, auto_reset=auto_reset,
                    aggregate_reward=auto_reset)
    pool = synthetic_pool(gc, 48, seed=4)
    n = 4000     # not a multiple of 32: the filler envs of the last warp must not be counted
    env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
    orc = OracleEnv(gc, n, n_threads=8)
    orc.upload_scenarios(pool)
    env.reset()
    orc.reset()
    rng = np.random.RandomState(5)
    want = np.zeros(abi.STAT_COUNT, np.float64)
    done_prev = np.zeros(n, bool)
    ep_steps = np.zeros(n, np.int64)
    for t in range(60 if auto_reset else 40):
        a = sample_actions(gc, rng, n, t)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        oo = orc.step(a)
        assert np.array_equal(done.cpu().numpy(), oo.done.astype(bool))
        ep_steps += 1
        if auto_reset:
            new = oo.done.astype(bool)
            v = _stats_from_outputs(np.zeros(n, bool), oo, orc.get_state())
            v[abi.STAT_RETURN_SUM] = oo.reward[new].astype(np.float64).sum()
            v[abi.STAT_LENGTH_SUM] = (ep_steps[new] * gc.c.frames_per_step).sum()
            ep_steps[new] = 0
            want += v
        else:
            want += _stats_from_outputs(done_prev, oo, orc.get_state())
            done_prev = oo.done.astype(bool).copy()
    got = env.stats().cpu().numpy()
    if auto_reset:
        assert want[abi.STAT_EPISODES] > 2 * n     # 26 steps per episode at most
    else:
        assert want[abi.STAT_EPISODES] == n        # max_steps = 250 frames: every env finished exactly once
    for k in (abi.STAT_EPISODES, abi.STAT_LENGTH_SUM, abi.STAT_CRASH, abi.STAT_SUCCESS, abi.STAT_TIMEOUT,
              abi.STAT_LEADER_CRASH):
        assert got[k] == want[k], (abi.STAT_NAMES[k], got[k], want[k])
    assert got[abi.STAT_RETURN_SUM] == pytest.approx(want[abi.STAT_RETURN_SUM], rel=1e-6 if auto_reset else 1e-9, abs=1e-6)
    assert got[abi.STAT_OVERFLOW] == 0
    # reset_after zeroes the vector
    env.stats(reset=True)
    assert float(env.stats().abs().sum()) == 0.0
    env.close()


def test_bench_workload_sample_matches_the_oracle_after_hundreds_of_steps():
    """The configuration bench.py times (BASELINE.json configs[2]: 65 536 envs, reference scenario pool, auto-reset,
    uniform random actions): envs [first, first + m) are re-simulated by the oracle with the same global env ids
    and the same actions; states are compared after 100, 200 and 300 steps -- i.e. across many in-step resets."""
    import torch
    from oracle_py import OracleEnv
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    import bench
    gc = bench.workload_config(auto_reset=True)
    pool, kind = bench.workload_pool(gc)
    n, first, m = 65536, 40960, 1024
    env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
    orc = OracleEnv(gc, m, env_id_base=first, n_threads=os.cpu_count() or 8)
    orc.upload_scenarios(pool)
    env.reset()
    orc.reset()
    _compare_states(env.get_state(first, m), orc.get_state(), gc, m, parity.RTOL)
    g = torch.Generator(device="cuda").manual_seed(99)
    lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
    bad, total, episodes = 0, 0, 0
    for t in range(300):
        a = (lo + (hi - lo) * torch.rand((n, 2), generator=g, device="cuda")).contiguous()
        obs, rew, done, info = env.step(a)
        oo = orc.step(a[first:first + m].cpu().numpy())
        episodes += int(oo.done.sum())
        if t % 20 == 19:
            assert np.array_equal(done[first:first + m].cpu().numpy(), oo.done.astype(bool)), "done differs at step %d" % t
            assert np.allclose(rew[first:first + m].cpu().numpy(), oo.reward, rtol=1e-5, atol=1e-5)
            bad += _ray_outliers(env.rays[first:first + m].cpu().numpy(), oo.rays)
            total += oo.rays.size
        if t + 1 in (100, 200, 300):
            _compare_states(env.get_state(first, m), orc.get_state(), gc, m, parity.RTOL)
    assert episodes > m // 2, "the sample saw too few in-step resets (%d)" % episodes
    assert bad <= 3, "%d of %d ray values outside tolerance" % (bad, total)
    env.close()


def test_device_rollout_matches_a_manual_loop_and_stays_on_the_device():
    """rollout.DeviceRollout: the trajectory it stores must be the one a hand-written loop over
    FtlBatchEnv.step produces from the same actions, the observation it feeds the policy is the fused sensorPrev matrix
    (WRP:203-221) of the raw sensor blocks, the stored actions are the policy's, and nothing in it lives on the host."""
    import torch
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    from continiousenvironment_follower_leader_b200.rollout import DeviceRollout
    from continiousenvironment_follower_leader_b200.wrappers import sensor_prev_observation
    kwargs = dict(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=True, max_steps=200)
    gc_f, gc_raw = GameConfig(fused_sensor_prev=True, **kwargs), GameConfig(**kwargs)
    pool = synthetic_pool(gc_f, 32, seed=6)
    n, T = 2048, 24
    for fused_policy, use_graphs in ((True, False), (False, True), (False, False)):
        ro = DeviceRollout(n, T, game_config=gc_f, scenario_pool=pool, seed=3, use_graphs=use_graphs,
                           fused_policy=fused_policy)
        assert (ro._fused is not None) == fused_policy
        # the fused kernel computes in bfloat16: ~3 digits on the pre-activation values
        atol_a, tol_v = (3e-2, 6e-2) if fused_policy else (2e-3, 1e-2)
        launches0 = ro.env.launch_count
        traj = ro.collect(explore=False)
        torch.cuda.synchronize()
        assert ro.env.launch_count - launches0 >= 3 * T
        for k, v in traj.items():
            assert v.is_cuda, k
        assert traj["obs"].shape == (T + 1, n, 240) and traj["actions"].shape == (T, n, 2)
        assert float(traj["obs"].min()) >= 0 and float(traj["obs"].max()) <= 1
        lo, hi = gc_f.action_bounds()
        assert bool((traj["actions"] >= torch.tensor(lo, device="cuda") - 1e-6).all())
        assert bool((traj["actions"] <= torch.tensor(hi, device="cuda") + 1e-6).all())
        assert int(traj["dones"].sum()) > 0          # max_steps = 200 frames: episodes end inside the window
        # the stored actions replayed by hand over the raw-sensor configuration
        env = FtlBatchEnv(n, game_config=gc_raw, scenario_pool=pool)
        env.reset()
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        with torch.no_grad():
            for t in range(T):
                obs = sensor_prev_observation(env).reshape(n, -1)
                assert torch.equal(obs.to(traj["obs"].dtype), traj["obs"][t]), "observation differs at step %d" % t
                act, val = ro.policy(obs, None)
                assert torch.allclose(act, traj["actions"][t], rtol=0, atol=atol_a * float(hi.max())), \
                    float((act - traj["actions"][t]).abs().max())
                assert torch.allclose(val, traj["values"][t], rtol=tol_v, atol=tol_v), float((val - traj["values"][t]).abs().max())
                _, rew, done, _ = env.step(traj["actions"][t].contiguous())
                assert torch.equal(rew, traj["rewards"][t]) and torch.equal(done, traj["dones"][t].bool())
        torch.backends.cuda.matmul.allow_tf32 = tf32
        adv, ret = ro.advantages()
        assert adv.shape == (T, n) and bool(torch.isfinite(adv).all())
        env.close()
        ro.close()


def _per_step_inputs_case():
    from continiousenvironment_follower_leader_b200.config import TEST_GAME_MANUAL_GAZEBO_KWARGS
    return dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, max_steps=700, auto_reset=True, random_frames_per_step=[1, 9],
                leader_speed_regime={0: [0.2, 1], 60: 1, 120: [0.5, 1], 200: 0.75, 260: [0.0, 0.5], 330: [0.4, 1]},
                leader_acceleration_regime={0: 0, 150: 0.03, 250: 0})


def run_per_step_inputs_case(make_env, n, steps, float_rtol):
    """FtlStepInputs on a batch: every env runs its own number of frames in every step (random_frames_per_step,
    ENV:939-940) and takes its list-valued speed-regime draws (ENV:1155-1156) from the caller; the stepper under test and
    the oracle receive the same inputs.  Shared by the CPU (host build) and GPU suites."""
    import warnings
    from oracle_py import OracleEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gc = GameConfig(**_per_step_inputs_case())
    assert gc.c.frames_per_step == 8 and gc.random_frames_per_step == (1, 9)
    pool = synthetic_pool(gc, 32, seed=9)
    sim, orc = make_env(gc, n), OracleEnv(gc, n, n_threads=8)
    sim.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    sim.reset(scenario_ids=ids)
    orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(23)
    bad, dones = 0, 0
    for t in range(steps):
        a = parity.sample_actions(gc, rng, n, t)
        frames = rng.randint(1, 9, size=n).astype(np.int32)
        draws = rng.random_sample((n, gc.c.frames_per_step))
        oc, oo = sim.step(a, frames=frames, regime_draws=draws), orc.step(a, frames=frames, regime_draws=draws)
        assert np.array_equal(oc.done, oo.done), "done differs at step %d" % t
        assert np.array_equal(oc.status, oo.status), "status differs at step %d" % t
        assert np.allclose(oc.reward, oo.reward, rtol=1e-5, atol=1e-5)
        assert np.allclose(oc.numerical_features, oo.numerical_features, rtol=max(float_rtol, 1e-12), atol=1e-4 if float_rtol else 0)
        bad += _ray_outliers(oc.rays, oo.rays)
        dones += int(oo.done.sum())
        if t % 10 == 9 or t == steps - 1:
            _compare_states(sim.get_state(), orc.get_state(), gc, n, float_rtol)
    assert bad <= 2 and dones > 0
    st = orc.get_state()
    assert len(np.unique(st.env["step_count"])) > 4    # the envs really ran different numbers of frames


def test_per_step_inputs_on_the_gpu():
    run_per_step_inputs_case(lambda gc, n: capi.HostEnv(gc, n, lib=capi.load()), 1024, 120, parity.RTOL)


def test_fused_policy_kernel_against_torch():
    """ftl_policy_mlp (csrc/ftl_policy.cu, bfloat16 mma) against the same MLP in torch float32 on bfloat16-rounded
    inputs and weights: what is left is the accumulation order and the bfloat16 rounding of the hidden activations."""
    import ctypes as C
    import torch
    from continiousenvironment_follower_leader_b200 import abi
    from continiousenvironment_follower_leader_b200.rollout import MlpPolicy
    L = capi.load()
    torch.manual_seed(0)
    # both kernels behind ftl_policy_mlp: tcgen05 / tensor memory (ftl_policy_tc.cu; obs_dim <= 256) and mma.sync (ftl_policy.cu)
    for impl, n, D, A in (("tcgen05", 1000, 240, 2), ("tcgen05", 131, 48, 1), ("tcgen05", 70000, 240, 2), ("tcgen05", 4096, 256, 7), ("tcgen05", 1, 240, 2), ("tcgen05", 129, 16, 3),
                          ("mma", 1000, 240, 2), ("mma", 131, 48, 1), ("mma", 4096, 288, 7), ("tcgen05", 4096, 288, 7)):
        os.environ["FTL_POLICY_IMPL"] = impl
        lo, hi = -np.arange(1, A + 1, dtype=np.float32), np.arange(1, A + 1, dtype=np.float32) * 2
        pol = MlpPolicy(D, lo, hi, seed=1).cuda()
        with torch.no_grad():
            for lin in (pol.body[0], pol.body[2], pol.head):
                lin.bias.normal_(0, 0.3)
        obs = torch.rand(n, D, device="cuda")
        noise = torch.randn(n, A, device="cuda")
        keep = {"w1": pol.body[0].weight.detach().to(torch.bfloat16).contiguous(), "b1": pol.body[0].bias.detach().contiguous(),
                "w2": pol.body[2].weight.detach().to(torch.bfloat16).contiguous(), "b2": pol.body[2].bias.detach().contiguous(),
                "w3": pol.head.weight.detach().to(torch.bfloat16).contiguous(), "b3": pol.head.bias.detach().contiguous(),
                "ns": pol.log_std.detach().exp().contiguous(), "mid": pol.act_mid.contiguous(), "half": pol.act_half.contiguous()}
        w = abi.FtlMlpWeights(*[keep[k].data_ptr() for k in ("w1", "b1", "w2", "b2", "w3", "b3", "ns", "mid", "half")], D, A)
        act, val = torch.zeros(n, A, device="cuda"), torch.zeros(n, device="cuda")
        capi.check(L, L.ftl_policy_mlp(C.byref(w), obs.data_ptr(), D, noise.data_ptr(), n, act.data_ptr(), val.data_ptr(), None),
                   "ftl_policy_mlp")
        torch.cuda.synchronize()
        with torch.no_grad():   # the same arithmetic in float32 on the rounded operands
            x = obs.to(torch.bfloat16).float()
            h = torch.tanh(x @ keep["w1"].float().T + keep["b1"]).to(torch.bfloat16).float()
            h = torch.tanh(h @ keep["w2"].float().T + keep["b2"]).to(torch.bfloat16).float()
            o = h @ keep["w3"].float().T + keep["b3"]
            want_a = keep["mid"] + keep["half"] * torch.tanh(o[:, :A] + noise * keep["ns"])
            want_v = o[:, A]
        assert torch.allclose(val, want_v, rtol=3e-3, atol=3e-3), (impl, n, D, A, float((val - want_v).abs().max()))
        assert torch.allclose(act, want_a, rtol=2e-3, atol=2e-3 * float(hi.max())), (impl, n, D, A, float((act - want_a).abs().max()))
    os.environ.pop("FTL_POLICY_IMPL", None)


def test_rgb_array_rasteriser_against_its_numpy_specification():
    """ftl_render / ftl_render_host (ENV:1196-1302) against tests/render_ref.py on envs in mid-episode: every pixel equal,
    except where a float32 distance sits on a boundary (a handful per frame) and inside the discs of the few trail points
    whose green-zone membership the step kernel leaves undecided."""
    import torch
    import render_ref
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=True)
    pool = synthetic_pool(gc, 16, seed=4)
    n = 64
    env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
    env.reset()
    rng = np.random.default_rng(0)
    lo, hi = gc.action_bounds()
    for _ in range(70):
        env.step(torch.as_tensor(rng.uniform(lo, hi, size=(n, 2)).astype(np.float32), device="cuda"))
    st = env.get_state()
    for scale in (4, 3):
        img = env.render(first=0, n=8, scale=scale).cpu().numpy()
        assert img.shape == (8, -(-gc.c.game_height // scale), -(-gc.c.game_width // scale), 3) and img.dtype == np.uint8
        for k in range(8):
            want, maybe_green = render_ref.render_env(gc.c, pool, st.env[k], st.trail[k], st.hist[k], st.corridor[k], scale)
            diff = (img[k] != want).any(axis=2)
            green_extra = diff & maybe_green & (img[k] == (0, 255, 0)).all(axis=2)
            bad = diff & ~green_extra
            assert bad.sum() <= 1e-3 * bad.size, "env %d scale %d: %d of %d pixels differ" % (k, scale, bad.sum(), bad.size)
            # what the picture is for: the follower's hit box is where the state says, in its colour
            r = st.env[k]["follower"]["rect"]
            cx, cy = (r[0] + r[2] // 2) // scale, (r[1] + r[3] // 2) // scale
            if 0 <= cx < img.shape[2] and 0 <= cy < img.shape[1] and r[2] >= 2 * scale and r[3] >= 2 * scale:
                assert tuple(img[k, cy, cx]) in ((255, 140, 0), (139, 69, 19), (80, 10, 10), (150, 120, 50), (255, 0, 0))
            assert (img[k] == 255).all(axis=2).mean() > 0.3      # mostly background
    env.close()
    # the host entry point behind gym_surface.Game.render("rgb_array"): full resolution, same picture as the device one
    host = capi.HostEnv(gc, 2, lib=capi.load())
    host.upload_scenarios(pool)
    host.reset(scenario_ids=np.array([1, 2], np.int32))
    for _ in range(5):
        host.step(rng.uniform(lo, hi, size=(2, 2)).astype(np.float32))
    full = host.render(0, 2, scale=1)
    assert full.shape == (2, gc.c.game_height, gc.c.game_width, 3)
    hs = host.get_state()
    want, maybe_green = render_ref.render_env(gc.c, pool, hs.env[0], hs.trail[0], hs.hist[0], hs.corridor[0], 1)
    bad = (full[0] != want).any(axis=2) & ~maybe_green
    assert bad.sum() <= 1e-3 * bad.size
    host.close()


def test_kin_pdl_option_changes_nothing_but_the_launch():
    """ftl_set_option(FTL_OPT_KIN_PDL): k_kin launched as a programmatic dependent of the kernel in front of it must give
    the same states and observations as plain stream order (it is off by default because it is slower, not because it
    differs)."""
    import torch
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=True, max_steps=300)
    pool = synthetic_pool(gc, 32, seed=9)
    n, steps = 4096, 40
    rng = np.random.default_rng(3)
    lo, hi = gc.action_bounds()
    acts = [torch.as_tensor(rng.uniform(lo, hi, size=(n, 2)).astype(np.float32), device="cuda") for _ in range(steps)]
    results = []
    for opt in (0, 1):
        env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
        capi.check(env._L, env._L.ftl_set_option(env._h, abi.OPT_KIN_PDL, opt), "ftl_set_option")
        env.reset()
        for a in acts:
            env.step_raw(a)
        torch.cuda.synchronize()
        st = env.get_state()
        results.append((st.env.tobytes(), env.rays.cpu().numpy().copy(), env.reward.cpu().numpy().copy()))
        assert env._L.ftl_set_option(env._h, 99, 1) != 0     # unknown options are refused
        env.close()
    assert results[0][0] == results[1][0]
    assert np.array_equal(results[0][1], results[1][1]) and np.array_equal(results[0][2], results[1][2])


def test_step_phase_options_split_a_step_without_changing_it():
    """ftl_set_option(FTL_OPT_STEP_PHASE): a step issued as its kinematics half (1) and its ray half (2) -- with and
    without the dependent-launch overlap (FTL_OPT_NO_OVERLAP) -- gives the states, rays and rewards of whole steps."""
    import torch
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=True, max_steps=300)
    pool = synthetic_pool(gc, 32, seed=9)
    n, steps = 4096, 40
    rng = np.random.default_rng(5)
    lo, hi = gc.action_bounds()
    acts = [torch.as_tensor(rng.uniform(lo, hi, size=(n, 2)).astype(np.float32), device="cuda") for _ in range(steps)]
    results = []
    for split, no_overlap in ((False, 0), (True, 0), (True, 1)):
        env = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
        capi.check(env._L, env._L.ftl_set_option(env._h, abi.OPT_NO_OVERLAP, no_overlap), "ftl_set_option")
        env.reset()
        for a in acts:
            if split:
                for phase in (1, 2):
                    capi.check(env._L, env._L.ftl_set_option(env._h, abi.OPT_STEP_PHASE, phase), "ftl_set_option")
                    env.step_raw(a)
            else:
                env.step_raw(a)
        torch.cuda.synchronize()
        capi.check(env._L, env._L.ftl_set_option(env._h, abi.OPT_STEP_PHASE, 0), "ftl_set_option")
        assert env._L.ftl_set_option(env._h, abi.OPT_STEP_PHASE, 3) != 0
        st = env.get_state()
        results.append((st.env.tobytes(), env.rays.cpu().numpy().copy(), env.reward.cpu().numpy().copy()))
        env.close()
    for r in results[1:]:
        assert r[0] == results[0][0]
        assert np.array_equal(r[1], results[0][1]) and np.array_equal(r[2], results[0][2])


def test_routes_through_three_finish_points_on_the_gpu():
    """multiple_end_points (ENV:471-482, 1552-1611): scenarios of the native generator with three chained D* legs (up to ~270
    waypoints, route_cap 512) on the CUDA path against the oracle."""
    from oracle_py import OracleEnv
    from continiousenvironment_follower_leader_b200 import scenario_gen
    gc = GameConfig(multiple_end_points=True, bear_number=1, follower_sensors=cfg3_sensors(), max_steps=1500)
    pool = scenario_gen.generate_pool_native(gc, [1, 2, 3, 4, 6, 8])
    assert gc.c.route_cap == 512 and int(pool.n_route.max()) > 128
    n, steps = 768, 220
    cuda, orc = capi.HostEnv(gc, n, lib=capi.load()), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    cuda.reset(scenario_ids=ids)
    orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(11)
    lo, hi = gc.action_bounds()
    bad = 0
    for t in range(steps):
        a = rng.uniform(lo, hi, size=(n, 2)).astype(np.float32)
        a[: n // 2] = (0.6 * hi[0], 0.0)      # half of the batch just drives on: its leaders get far along their routes
        oc, oo = cuda.step(a), orc.step(a)
        assert np.array_equal(oc.done, oo.done), "done differs at step %d" % t
        assert np.array_equal(oc.status, oo.status), "status differs at step %d" % t
        assert np.array_equal(oc.leader_target, oo.leader_target)
        assert np.allclose(oc.numerical_features, oo.numerical_features, rtol=parity.RTOL, atol=1e-4)
        bad += _ray_outliers(oc.rays, oo.rays)
        if t % 20 == 19 or t == steps - 1:
            _compare_states(cuda.get_state(), orc.get_state(), gc, n, parity.RTOL)
    assert bad <= 2
    assert int(orc.get_state().env["cur_target_id"].max()) > 45      # leaders are well into their routes (56 here)
    cuda.close()
