"""GPU parity tests: libftl.so (CUDA, through the C-ABI) against the golden reference traces and the
CPU oracle.  Run on the B200 box with `pytest -m gpu`.

Bars (BASELINE.json north_star): collision/done/step-count and every other integer bit-exact;
poses, rewards and ray distances within 1e-4 relative.  Ray values may additionally differ where a
ray grazes a corner (the hit/no-hit predicate is discontinuous there); those are counted and must stay
below 1e-4 of all ray values.
"""
import numpy as np
import pytest

import parity
from continiousenvironment_follower_leader_b200 import abi, capi
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors, TEST_GAME_MANUAL_GAZEBO_KWARGS
from continiousenvironment_follower_leader_b200.scenario import synthetic_pool

pytestmark = pytest.mark.gpu

FILES = parity.golden_files()


def _cuda_env(gc, n, **kw):
    return capi.HostEnv(gc, n, lib=capi.load(), **kw)   # raises if libftl.so is missing: no fallback


@pytest.mark.parametrize("path", FILES, ids=[p.split("/")[-1][:-4] for p in FILES])
def test_cuda_reproduces_reference_trace(path):
    d, meta = parity.load_trace(path)
    gc = parity.config_for(meta, _route_len=len(d["scen_route"]), _n_static=len(d["scen_static_rects"]))
    budget = 0     # hit / no-hit decisions are the reference's by construction (float64 fallback), values within 1e-4
    for idx in (0, 32):
        # a fresh handle per replay: like the reference's env object, a handle never restores the keys of
        # leader_acceleration_regime it has consumed (ENV:1170), so a second episode on it is a different episode
        env = _cuda_env(gc, 33)   # 33 copies: more than a warp, all must agree
        env.upload_scenarios(parity.pool_for(d, gc))
        T, outliers = parity.replay(env, d, gc, env_index=idx, float_rtol=parity.RTOL, ray_rtol=parity.RTOL,
                                    ray_outlier_budget=budget)
        assert T == meta["n_env_steps"]
        parity.check_final_arrays(env.get_state(), d, gc, env_index=idx)
        env.close()


def _compare_states(a, b, gc, n, float_rtol):
    """a: cuda HostState, b: oracle HostState.  Integers exact, floats within tolerance."""
    ea, eb = a.env, b.env
    int_fields = ["scenario_id", "cur_target_id", "leader_finished", "step_count", "finish_timer", "done", "crash",
                  "is_in_box", "is_on_trace", "too_close", "mission_status", "agent_status", "leader_status",
                  "trail_len", "saving_counter", "ring_tail", "ring_head", "hist_f64_end", "episode_count", "overflow",
                  "accel_consumed"]
    for f in int_fields:
        assert np.array_equal(ea[f], eb[f]), "%s differs at envs %s" % (f, np.nonzero(ea[f] != eb[f])[0][:8])
    robots = ["follower", "leader"]
    for r in robots:
        for f in ("rect", "rot_dir", "des_rot_dir"):
            assert np.array_equal(ea[r][f], eb[r][f]), "%s.%s differs" % (r, f)
        for f in ("pos", "dir", "speed", "rot_speed", "des_speed", "des_rot_speed"):
            x, y = ea[r][f].astype(np.float64), eb[r][f].astype(np.float64)
            assert np.all(np.abs(x - y) <= float_rtol * np.maximum(1, np.abs(y))), "%s.%s differs" % (r, f)
    nb = gc.c.n_bears
    if nb:
        assert np.array_equal(ea["bear"]["rect"][:, :nb], eb["bear"]["rect"][:, :nb])
        assert np.array_equal(ea["bear_index"][:, :nb], eb["bear_index"][:, :nb])
        x, y = ea["bear"]["pos"][:, :nb].astype(np.float64), eb["bear"]["pos"][:, :nb].astype(np.float64)
        assert np.all(np.abs(x - y) <= float_rtol * np.maximum(1, np.abs(y)))
    for f in ("overall_reward", "accumulated_penalty", "last_reward"):
        assert np.all(np.abs(ea[f] - eb[f]) <= float_rtol * np.maximum(1, np.abs(eb[f]))), f


def _ray_outliers(got, want, rtol=parity.RTOL):
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    return int(np.sum(err > rtol * np.maximum(1.0, np.abs(want))))


@pytest.mark.parametrize("name,kwargs,n,steps", [
    ("cfg2", dict(add_obstacles=False, add_bear=False,
                  follower_sensors={"LeaderPositionsTracker_v2": cfg3_sensors()["LeaderPositionsTracker_v2"]}), 4096, 40),
    ("cfg3", dict(bear_number=1, follower_sensors=cfg3_sensors()), 2048, 60),
    ("cfg3_3bears_discrete", dict(bear_number=3, discrete_action_space=True, follower_sensors=cfg3_sensors(24, 20, 3)), 512, 40),
    ("cfg4_gazebo_hardcore", dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, max_steps=900, auto_reset=True), 1024, 120),
    ("cfg5_rays_360", dict(bear_number=2, frames_per_step=3, follower_sensors=cfg3_sensors(72, 360, 4)), 256, 30),
])
def test_cuda_matches_oracle_on_seeded_batch(name, kwargs, n, steps):
    from oracle_py import OracleEnv
    gc = GameConfig(**kwargs)
    pool = synthetic_pool(gc, 64, seed=1)
    cuda, orc = _cuda_env(gc, n), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    oc, oo = cuda.reset(scenario_ids=ids), orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(7)
    bounds = gc.action_bounds()
    total_rays, bad_rays = 0, 0
    for t in range(steps):
        if gc.discrete_action_space:
            a = rng.randint(0, 5, size=n).astype(np.int32)
        else:
            a = rng.uniform(bounds[0], bounds[1], size=(n, 2)).astype(np.float32)
        oc, oo = cuda.step(a), orc.step(a)
        assert np.array_equal(oc.done, oo.done), "done differs at step %d" % t
        assert np.array_equal(oc.status, oo.status), "status differs at step %d" % t
        assert np.allclose(oc.reward, oo.reward, rtol=1e-5, atol=1e-6), "reward differs at step %d" % t
        assert np.array_equal(oc.leader_target, oo.leader_target)
        assert np.allclose(oc.numerical_features, oo.numerical_features, rtol=parity.RTOL, atol=1e-4)
        if gc.rays_per_env:
            bad_rays += _ray_outliers(oc.rays, oo.rays)
            total_rays += oc.rays.size
        if t % 10 == 9 or t == steps - 1:
            _compare_states(cuda.get_state(), orc.get_state(), gc, n, parity.RTOL)
    if total_rays:
        assert bad_rays <= 2, "%d of %d ray values outside tolerance" % (bad_rays, total_rays)


def test_auto_reset_matches_oracle():
    from oracle_py import OracleEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), max_steps=300, auto_reset=True)
    pool = synthetic_pool(gc, 32, seed=3)
    n = 256
    cuda, orc = _cuda_env(gc, n), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    cuda.reset()
    orc.reset()
    rng = np.random.RandomState(11)
    lo, hi = gc.action_bounds()
    dones = 0
    for t in range(80):
        a = rng.uniform(lo, hi, size=(n, 2)).astype(np.float32)
        oc, oo = cuda.step(a), orc.step(a)
        assert np.array_equal(oc.done, oo.done)
        assert np.allclose(oc.numerical_features, oo.numerical_features, rtol=parity.RTOL, atol=1e-4)
        dones += int(oc.done.sum())
    assert dones > n            # every env finished at least once (max_steps=300 frames = 30 steps)
    sa, sb = cuda.get_state(), orc.get_state()
    assert np.array_equal(sa.env["episode_count"], sb.env["episode_count"])
    assert np.array_equal(sa.env["scenario_id"], sb.env["scenario_id"])
    _compare_states(sa, sb, gc, n, parity.RTOL)


def test_teacher_forced_state_injection():
    """set_state(oracle state) then one step: single-step parity independent of trajectory history."""
    from oracle_py import OracleEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors())
    pool = synthetic_pool(gc, 16, seed=5)
    n = 128
    cuda, orc = _cuda_env(gc, n), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    orc.reset()
    cuda.reset()
    rng = np.random.RandomState(2)
    lo, hi = gc.action_bounds()
    for t in range(25):
        orc.step(rng.uniform(lo, hi, size=(n, 2)).astype(np.float32))
    cuda.set_state(orc.get_state())
    for t in range(5):
        a = rng.uniform(lo, hi, size=(n, 2)).astype(np.float32)
        oc, oo = cuda.step(a), orc.step(a)
        assert np.array_equal(oc.done, oo.done)
        assert _ray_outliers(oc.rays, oo.rays) <= 2
        _compare_states(cuda.get_state(), orc.get_state(), gc, n, parity.RTOL)


def test_full_size_properties():
    """BASELINE.json configs[2] size (65536 envs): determinism, slice invariance, output ranges."""
    import torch
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=True)
    pool = synthetic_pool(gc, 128, seed=0)
    n = 65536

    def run(n_envs, base):
        env = FtlBatchEnv(n_envs, game_config=gc, scenario_pool=pool, env_id_base=base)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(1)
        lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
        acts = lo + (hi - lo) * torch.rand((20, n, 2), generator=g, device="cuda")
        for t in range(20):
            obs, rew, done, info = env.step(acts[t, base:base + n_envs].contiguous())
        out = {k: v.clone() for k, v in obs.items()}
        out["reward"], out["done"] = rew.clone(), done.clone()
        st = env.get_state(0, min(n_envs, 512))
        env.close()
        return out, st

    a, sa = run(n, 0)
    b, sb = run(n, 0)
    for k in a:
        assert torch.equal(a[k], b[k]), "non-deterministic: " + k
    c, sc = run(4096, 8192)     # a slice of the batch run on its own must give the same results
    for k in a:
        assert torch.equal(a[k][8192:8192 + 4096], c[k]), "batch composition changed results: " + k
    for name, off, h, w in gc.ray_layout():
        L = [r.laser_length for r in gc.c.ray[:gc.c.n_ray_sensors]][[x[0] for x in gc.ray_layout()].index(name)]
        v = a[name]
        assert float(v.min()) > 0 and float(v.max()) <= L * (1 + 1e-6)
    nf = a["numerical_features"]
    assert torch.isfinite(nf).all()
    assert int(sa.env["overflow"].max()) == 0


def test_device_and_host_entry_points_agree():
    import torch
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors())
    pool = synthetic_pool(gc, 8, seed=2)
    n = 300
    dev = FtlBatchEnv(n, game_config=gc, scenario_pool=pool)
    host = _cuda_env(gc, n)
    host.upload_scenarios(pool)
    dev.reset()
    host.reset()
    rng = np.random.RandomState(0)
    lo, hi = gc.action_bounds()
    for t in range(12):
        a = rng.uniform(lo, hi, size=(n, 2)).astype(np.float32)
        obs, rew, done, info = dev.step(torch.from_numpy(a).cuda())
        oh = host.step(a)
        assert np.array_equal(obs["numerical_features"].cpu().numpy(), oh.numerical_features)
        assert np.array_equal(dev.rays.cpu().numpy(), oh.rays)
        assert np.array_equal(rew.cpu().numpy(), oh.reward)
    assert dev.launch_count >= 24


def test_pipelined_halves_equal_the_whole_batch():
    """Two half-batch handles on their own streams, driven with step_begin / step_wait (ftl_step_host_begin/_wait),
    must deliver exactly what one synchronous whole-batch handle delivers -- auto-reset included (env ids are global)."""
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=True, max_steps=150)
    pool = synthetic_pool(gc, 8, seed=5)
    n, half = 9000, 4480   # large enough for the chunked ray launches of the host path (>= 8192 envs)
    whole = _cuda_env(gc, n, pinned=True)
    parts = [_cuda_env(gc, m, env_id_base=first, pinned=True, own_stream=True) for first, m in ((0, half), (half, n - half))]
    for e in [whole] + parts:
        e.upload_scenarios(pool)
        e.reset()
    rng = np.random.RandomState(3)
    lo, hi = gc.action_bounds()
    acts = [rng.uniform(lo, hi, size=(n, 2)).astype(np.float32) for _ in range(25)]
    A, B = parts
    A.step_begin(acts[0][:half])
    with pytest.raises(capi.FtlError):
        A.step_begin(acts[0][:half])           # one step in flight per handle
    for t, a in enumerate(acts):
        B.step_begin(a[half:])
        oa = A.step_wait()
        ow = whole.step(a)
        for f in ("numerical_features", "leader_target", "rays", "reward", "done", "status"):
            assert np.array_equal(getattr(oa, f), getattr(ow, f)[:half]), (f, t)
        if t + 1 < len(acts):
            A.step_begin(acts[t + 1][:half])
        ob = B.step_wait()
        for f in ("numerical_features", "leader_target", "rays", "reward", "done", "status"):
            assert np.array_equal(getattr(ob, f), getattr(ow, f)[half:]), (f, t)
    assert whole.out.done.sum() >= 0
    for e in [whole] + parts:
        e.close()


def test_grazing_rays_resolve_like_the_reference_on_the_gpu():
    """Followers parked on integer coordinates with axis-aligned headings: many rays run exactly through corners and
    along edges.  k_rays records those pairs, k_rays_exact must reproduce the reference's strict ccw decisions."""
    from oracle_py import OracleEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(12, 36, 5))
    pool = synthetic_pool(gc, 8, seed=2)
    n = 256
    cuda, orc = _cuda_env(gc, n), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    cuda.reset(scenario_ids=ids)
    orc.reset(scenario_ids=ids)
    st = orc.get_state()
    rng = np.random.RandomState(3)
    for i in range(n):
        rects = pool.static_rects[ids[i], 2:pool.n_static[ids[i]]]
        x, y, w, h = rects[rng.randint(len(rects))]
        dx, dy = [(-60, 0), (-60, 25), (-60, h), (w + 40, -30), (w // 2, -50), (-40, -40)][i % 6]
        st.env["follower"]["pos"][i] = (x + dx, y + dy)
        st.env["follower"]["dir"][i] = 45.0 * (i % 8)
        st.env["follower"]["speed"][i] = 0.0
        st.env["follower"]["rot_speed"][i] = 0.0
    orc.set_state(st)
    cuda.set_state(st)
    zero = np.zeros((n, 2), np.float32)
    for t in range(4):
        a, b = cuda.step(zero), orc.step(zero)
        assert np.array_equal(a.numerical_features, b.numerical_features)
        assert _ray_outliers(a.rays, b.rays, rtol=1e-5) == 0


def test_gym_surface_on_the_gpu_replays_a_reference_trace():
    """gym_surface.Game (single-env view, batch of one) through libftl.so: reset on the exported scenario, 4-tuple
    steps, numerical features bit-equal to the reference trace, rays within 1e-4, same reward/done/info."""
    from continiousenvironment_follower_leader_b200 import gym_surface as gs, scenario_gen, wrappers
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/cfg3_seed5_follow.npz")
    env = gs.make("Test-Cont-Env-Auto-v0", **meta["kwargs"])
    wrapped = wrappers.ContinuousObserveModifier_sensorPrev(env, max_prev_obs=5)
    sc = scenario_gen.Scenario()
    sc.static_rects = [tuple(r) for r in d["scen_static_rects"]]
    sc.route = [tuple(p) for p in d["scen_route"]]
    sc.leader_pos, sc.leader_dir = d["scen_leader_pos"], float(d["scen_leader_dir"])
    sc.follower_pos, sc.follower_dir = d["scen_follower_pos"], float(d["scen_follower_dir"])
    sc.found_target_point = True
    obs = env.reset(scenario=sc)
    assert np.allclose(obs["numerical_features"], d["t_nf"][0], rtol=parity.RTOL)
    for t, a in enumerate(d["actions"][:260]):
        obs, reward, done, info = env.step(a)
        ints = d["t_ints"][t + 1]
        assert np.allclose(obs["numerical_features"], d["t_nf"][t + 1], rtol=parity.RTOL, atol=1e-4)
        assert abs(reward - d["t_floats"][t + 1][0]) < 1e-6 and done == bool(ints[4])
        assert info["mission_status"] == abi.MISSION_STATUS[ints[10]] and info["agent_status"] == abi.AGENT_STATUS[ints[11]]
        got = np.concatenate([obs[n].reshape(-1) for n in meta["ray_names"]])
        assert np.allclose(got, d["t_rays"][t + 1], rtol=parity.RTOL)
        feats = wrapped.observation(obs)
        assert feats.shape == (5, 48) and feats.min() >= 0 and feats.max() <= 1
    assert env.step_count == 2600


@pytest.mark.gpu
def test_fused_sensor_prev_output_on_the_gpu():
    """fused_sensor_prev=True on the device (both entry points) against the wrapper's arithmetic applied to the
    oracle's raw sensor output (WRP:203-221); the exact pass patches cells on the normalised scale."""
    from oracle_py import OracleEnv
    kwargs = dict(bear_number=1, follower_sensors=cfg3_sensors())
    n, steps = 2048, 40
    gc, gc_raw = GameConfig(fused_sensor_prev=True, **kwargs), GameConfig(**kwargs)
    pool = synthetic_pool(gc, 64, seed=1)
    cuda, orc = _cuda_env(gc, n), OracleEnv(gc_raw, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    oc, oo = cuda.reset(scenario_ids=ids), orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(7)
    bounds = gc.action_bounds()
    bad = 0
    for t in range(steps):
        want = parity.sensor_prev_expected(gc_raw, oo.rays)
        got = oc.rays.reshape(want.shape)
        bad += int(np.sum(np.abs(got - want) > parity.RTOL))
        a = rng.uniform(bounds[0], bounds[1], size=(n, 2)).astype(np.float32)
        oc, oo = cuda.step(a), orc.step(a)
    assert bad <= 2, "%d fused cells differ" % bad

    import torch
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    from continiousenvironment_follower_leader_b200.wrappers import sensor_prev_observation
    fused, raw = FtlBatchEnv(512, game_config=gc, scenario_pool=pool), FtlBatchEnv(512, game_config=gc_raw, scenario_pool=pool)
    ids_t = torch.arange(512, dtype=torch.int32, device="cuda") % pool.n
    fused.reset(scenario_ids=ids_t)
    raw.reset(scenario_ids=ids_t)
    for t in range(10):
        a = torch.from_numpy(rng.uniform(bounds[0], bounds[1], size=(512, 2)).astype(np.float32)).cuda()
        obs_f, *_ = fused.step(a)
        raw.step(a)
        assert torch.equal(obs_f["sensor_prev"], sensor_prev_observation(raw))


def _scenario_of(d):
    from continiousenvironment_follower_leader_b200 import scenario_gen
    sc = scenario_gen.Scenario()
    sc.static_rects = [tuple(r) for r in d["scen_static_rects"]]
    sc.route = [tuple(p) for p in d["scen_route"]]
    sc.leader_pos, sc.leader_dir = d["scen_leader_pos"], float(d["scen_leader_dir"])
    sc.follower_pos, sc.follower_dir = d["scen_follower_pos"], float(d["scen_follower_dir"])
    sc.found_target_point = True
    return sc


@pytest.mark.gpu
def test_gym_surface_returns_the_flat_sensors_like_the_reference():
    """LeaderCorridor_lasers_v2 / LeaderCorridor_lasers / FollowerInfo / LeaderTrackDetector_vector through the
    single-env view: same dict keys and shapes as the reference's use_sensors (CLS:255-288), values of the trace."""
    from continiousenvironment_follower_leader_b200 import gym_surface as gs
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/flat_sensors_seed9.npz")
    env = gs.make("Test-Cont-Env-Auto-v0", **meta["kwargs"])
    obs = env.reset(scenario=_scenario_of(d))
    for t, a in enumerate(d["actions"][:120]):
        obs, reward, done, info = env.step(a)
        assert obs["LeaderCorridor_lasers_v2"].shape == (24,) and obs["LeaderCorridor_lasers"].shape == (7,)
        assert obs["FollowerInfo"].shape == (2,) and obs["LeaderTrackDetector_vector"].shape == (12, 2)
        got = np.concatenate([obs[n].reshape(-1) for n in meta["ray_names"]])
        assert np.allclose(got, d["t_rays"][t + 1], rtol=parity.RTOL)
        assert np.allclose(obs["FollowerInfo"], d["t_follower_info"][t + 1], rtol=parity.RTOL, atol=1e-6)
        assert np.allclose(obs["LeaderTrackDetector_vector"], d["t_track_vectors"][t + 1], atol=0.15 * parity.RTOL * 1500)


@pytest.mark.gpu
def test_flat_sensors_match_the_oracle_on_a_batch():
    from oracle_py import OracleEnv
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/flat_sensors_old_seed13.npz")
    kwargs = dict(meta["kwargs"], auto_reset=True, max_steps=400)
    n, steps = 1024, 90
    gc = GameConfig(**kwargs)
    pool = synthetic_pool(gc, 64, seed=5)
    cuda, orc = _cuda_env(gc, n), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    oc, oo = cuda.reset(scenario_ids=ids), orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(3)
    bounds = gc.action_bounds()
    bad = 0
    for t in range(steps):
        assert np.allclose(oc.track_vectors, oo.track_vectors, atol=0.15), "track vectors differ at step %d" % t
        bad += _ray_outliers(oc.rays, oo.rays)
        a = rng.uniform(bounds[0], bounds[1], size=(n, 2)).astype(np.float32)
        a[: n // 2, 0], a[: n // 2, 1] = bounds[1][0], 0.0
        oc, oo = cuda.step(a), orc.step(a)
        assert np.array_equal(oc.done, oo.done) and np.array_equal(oc.status, oo.status)
    assert bad <= 2


def _radar_outliers(got, want):
    """Sectors whose value differs by more than 1e-4 relative: a history point within rounding of a sector boundary
    may fall into the neighbouring sector (arccos of float64 quotients; CUDA's libm differs from glibc in the last bit)."""
    return int(np.sum(np.abs(got.astype(np.float64) - want) > parity.RTOL * np.maximum(1.0, np.abs(want))))


@pytest.mark.gpu
@pytest.mark.parametrize("trace", ["radar_old_seed17", "radar_new_seed19", "radar_near_seed21"])
def test_radar_matches_the_oracle_on_a_batch(trace):
    """LeaderTrackDetector_radar (SEN:394-461) through the C-ABI against the oracle, which reproduces the reference's
    radar bit for bit on the golden traces."""
    from oracle_py import OracleEnv
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/%s.npz" % trace)
    kwargs = dict(meta["kwargs"], auto_reset=True, max_steps=400)
    n, steps = 1024, 80
    gc = GameConfig(**kwargs)
    pool = synthetic_pool(gc, 64, seed=6)
    cuda, orc = _cuda_env(gc, n), OracleEnv(gc, n, n_threads=8)
    cuda.upload_scenarios(pool)
    orc.upload_scenarios(pool)
    ids = (np.arange(n) % pool.n).astype(np.int32)
    oc, oo = cuda.reset(scenario_ids=ids), orc.reset(scenario_ids=ids)
    rng = np.random.RandomState(4)
    bounds = gc.action_bounds()
    bad, total, seen = 0, 0, 0
    for t in range(steps):
        assert oc.radar.shape == (n, gc.c.radar_sectors)
        bad += _radar_outliers(oc.radar, oo.radar)
        total += oc.radar.size
        seen += int((oo.radar > 0).sum())
        a = rng.uniform(bounds[0], bounds[1], size=(n, 2)).astype(np.float32)
        a[: n // 2, 0], a[: n // 2, 1] = bounds[1][0], 0.0
        oc, oo = cuda.step(a), orc.step(a)
        assert np.array_equal(oc.done, oo.done) and np.array_equal(oc.status, oo.status)
    assert seen > 1000, "the radar never saw the trail: the test is vacuous"
    assert bad <= max(4, 2e-5 * total), "%d of %d radar sectors outside tolerance" % (bad, total)


@pytest.mark.gpu
def test_gym_surface_returns_the_radar_like_the_reference():
    from continiousenvironment_follower_leader_b200 import gym_surface as gs
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/radar_old_seed17.npz")
    env = gs.make("Test-Cont-Env-Auto-v0", **meta["kwargs"])
    obs = env.reset(scenario=_scenario_of(d))
    assert _radar_outliers(obs["LeaderTrackDetector_radar"], d["t_radar"][0]) == 0
    bad = 0
    for t, a in enumerate(d["actions"][:150]):
        obs, reward, done, info = env.step(a)
        assert obs["LeaderTrackDetector_radar"].shape == (12,)
        bad += _radar_outliers(obs["LeaderTrackDetector_radar"], d["t_radar"][t + 1])
    assert bad <= 1


@pytest.mark.gpu
def test_overlapped_ray_kernel_with_a_multi_wave_step_kernel():
    """100 000 envs are more blocks of k_step than fit the GPU at once, so the programmatic dependent launch of k_rays
     starts while later waves of k_step are still queued: every ray warp must still see the state its
    env's step published.  A slice of the big batch run on its own (same global env ids) is the reference."""
    import torch
    from continiousenvironment_follower_leader_b200.batch_env import FtlBatchEnv
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), auto_reset=True, max_steps=150)
    pool = synthetic_pool(gc, 96, seed=3)
    n, first, m = 100000, 61440, 2048

    def run(n_envs, base):
        env = FtlBatchEnv(n_envs, game_config=gc, scenario_pool=pool, env_id_base=base)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(7)
        lo, hi = [torch.tensor(x, device="cuda") for x in gc.action_bounds()]
        acts = lo + (hi - lo) * torch.rand((30, n, 2), generator=g, device="cuda")
        rays, feats = [], []
        for t in range(30):
            obs, rew, done, info = env.step(acts[t, base:base + n_envs].contiguous())
            rays.append(env.rays.clone())
            feats.append(obs["numerical_features"].clone())
        env.close()
        return rays, feats

    big_r, big_f = run(n, 0)
    small_r, small_f = run(m, first)
    for t in range(30):
        assert torch.equal(big_f[t][first:first + m], small_f[t]), "state differs at step %d" % t
        assert torch.equal(big_r[t][first:first + m], small_r[t]), "rays differ at step %d" % t
